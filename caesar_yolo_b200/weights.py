"""YOLOv8 architecture table, seeded random-init weights (ultralytics state-dict key names) and the weight
file format accepted by `--weights`.

ultralytics is not installed in this environment, so real `.pt` checkpoints (which unpickle ultralytics classes)
cannot be loaded; the `--weights` file of this build is a plain torch-saved dict
{'format': 'caesar_yolo_b200-weights-v1', 'variant', 'nc', 'names', 'state_dict'} whose state_dict uses the
ultralytics key names (model.0.conv.weight, model.0.bn.running_mean, ..., model.22.cv3.2.2.bias).
"""
import json
import math
import os

import torch

CLASS_NAMES = {0: 'spurious', 1: 'compact', 2: 'extended', 3: 'extended-multisland', 4: 'flagged'}
SCALES = {'n': (0.33, 0.25, 1024), 's': (0.33, 0.50, 1024), 'm': (0.67, 0.75, 768), 'l': (1.00, 1.00, 512),
          'x': (1.00, 1.25, 512)}
FORMAT = 'caesar_yolo_b200-weights-v1'
_CALIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'init_calibration.json')


def make_divisible(x, d):
    return int(math.ceil(x / d) * d)


def arch(variant):
    """Channel widths / repeats of yolov8{n,s,m,l,x}.yaml (ultralytics cfg/models/v8/yolov8.yaml)."""
    depth, width, maxc = SCALES[variant]
    ch = lambda c: make_divisible(min(c, maxc) * width, 8)
    rep = lambda n: max(round(n * depth), 1)
    return dict(c1=ch(64), c2=ch(128), c3=ch(256), c4=ch(512), c5=ch(1024), n2=rep(3), n4=rep(6), n6=rep(6),
                n8=rep(3), nh=rep(3))


def conv_bn_layers(variant, nc=5):
    """Ordered list of (prefix, cin, cout, k) for every Conv(+BN+SiLU) module, in forward order."""
    a = arch(variant)
    c1, c2, c3, c4, c5 = a['c1'], a['c2'], a['c3'], a['c4'], a['c5']
    L = []

    def c2f(p, cin, cout, n):
        c = cout // 2
        L.append((p + '.cv1', cin, 2 * c, 1))
        for i in range(n):
            L.append(('%s.m.%d.cv1' % (p, i), c, c, 3))
            L.append(('%s.m.%d.cv2' % (p, i), c, c, 3))
        L.append((p + '.cv2', (2 + n) * c, cout, 1))

    L.append(('model.0', 3, c1, 3))
    L.append(('model.1', c1, c2, 3))
    c2f('model.2', c2, c2, a['n2'])
    L.append(('model.3', c2, c3, 3))
    c2f('model.4', c3, c3, a['n4'])
    L.append(('model.5', c3, c4, 3))
    c2f('model.6', c4, c4, a['n6'])
    L.append(('model.7', c4, c5, 3))
    c2f('model.8', c5, c5, a['n8'])
    L.append(('model.9.cv1', c5, c5 // 2, 1))
    L.append(('model.9.cv2', c5 * 2, c5, 1))
    c2f('model.12', c5 + c4, c4, a['nh'])
    c2f('model.15', c4 + c3, c3, a['nh'])
    L.append(('model.16', c3, c3, 3))
    c2f('model.18', c3 + c4, c4, a['nh'])
    L.append(('model.19', c4, c4, 3))
    c2f('model.21', c4 + c5, c5, a['nh'])
    cb = max(16, c3 // 4, 64)
    cc = max(c3, min(nc, 100))
    for l, c in enumerate((c3, c4, c5)):
        L.append(('model.22.cv2.%d.0' % l, c, cb, 3))
        L.append(('model.22.cv2.%d.1' % l, cb, cb, 3))
        L.append(('model.22.cv3.%d.0' % l, c, cc, 3))
        L.append(('model.22.cv3.%d.1' % l, cc, cc, 3))
    return L, cb, cc


def _load_calibration(variant, seed):
    if os.path.exists(_CALIB_PATH):
        with open(_CALIB_PATH) as f:
            t = json.load(f)
        return t.get('%s:%d' % (variant, seed))
    return None


def make_random_weights(variant='n', nc=5, seed=0, cls_bias=-3.0, calibration='auto'):
    """Seeded random-init YOLOv8 (our own init; the reference has none).  Conv weights ~ N(0, 1/fan_in); the BN
    running statistics come from init_calibration.json (per-layer scalar mean/var of the pre-BN activation on a
    synthetic preprocessed tile, produced by tools/calibrate_init.py for the (variant, seed) pairs it lists; other
    seeds fall back to mean 0 / var 1) so that every layer works at unit scale and detections depend on the image.  The class-branch bias sets the candidate density; a linear DFL bias keeps
    boxes a few cells wide."""
    calib = _load_calibration(variant, seed) if calibration == 'auto' else calibration
    g = torch.Generator().manual_seed(seed)
    layers, cb, cc = conv_bn_layers(variant, nc)
    sd = {}
    for (p, cin, cout, k) in layers:
        fan_in = cin * k * k
        sd[p + '.conv.weight'] = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(fan_in)
        mu, var = (calib[p] if calib and p in calib else (0.0, 1.0))
        sd[p + '.bn.weight'] = 1.0 + 0.1 * torch.randn(cout, generator=g)
        sd[p + '.bn.bias'] = 0.1 * torch.randn(cout, generator=g)
        sd[p + '.bn.running_mean'] = torch.full((cout,), float(mu))
        sd[p + '.bn.running_var'] = float(var) * (0.8 + 0.4 * torch.rand(cout, generator=g))
    for l in range(3):
        sd['model.22.cv2.%d.2.weight' % l] = torch.randn(64, cb, 1, 1, generator=g) * (1.5 / math.sqrt(cb))
        sd['model.22.cv2.%d.2.bias' % l] = (-0.6 * torch.arange(16, dtype=torch.float32)).repeat(4) + 1.0
        sd['model.22.cv3.%d.2.weight' % l] = torch.randn(nc, cc, 1, 1, generator=g) * (2.5 / math.sqrt(cc))
        sd['model.22.cv3.%d.2.bias' % l] = torch.full((nc,), float(cls_bias))
    sd['model.22.dfl.conv.weight'] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return {'format': FORMAT, 'variant': variant, 'nc': nc, 'names': dict(CLASS_NAMES), 'state_dict': sd}


def save_weights(weights, path):
    torch.save(weights, path)


def load_weights(path):
    w = torch.load(path, map_location='cpu', weights_only=False)
    if not isinstance(w, dict) or w.get('format') != FORMAT:
        raise ValueError("%s is not a %s file (ultralytics .pt checkpoints need the ultralytics package, "
                         "which is not available in this environment)" % (path, FORMAT))
    return w
