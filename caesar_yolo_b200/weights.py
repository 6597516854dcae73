"""YOLOv8 architecture table, seeded random-init weights (ultralytics state-dict key names) and the weight
file format accepted by `--weights`.

`--weights` (scripts/run.py:347, `YOLO(weights_path)`) accepts
  * an ultralytics YOLOv8 detection checkpoint (`.pt`: a pickled `{'model': DetectionModel, 'ema': ..., ...}`).  The
    ultralytics package is NOT needed: `load_ultralytics_checkpoint` unpickles with stand-in classes for every module
    that cannot be imported and reads the tensors out of the module tree (variant from the stem width, nc from the
    class branch, names from `model.names`); Conv+BN already fused (`model.fuse()`) checkpoints are accepted too;
  * this build's own file: a plain torch-saved dict
    {'format': 'caesar_yolo_b200-weights-v1', 'variant', 'nc', 'names', 'state_dict'} whose state_dict uses the
    ultralytics key names (model.0.conv.weight, model.0.bn.running_mean, ..., model.22.cv3.2.2.bias).
"""
import json
import math
import os
import pickle

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


# ------------------------------------------------------------------------------------------------ YOLO11
# ultralytics cfg/models/11/yolo11.yaml (the reference's README ships yolo11 weights next to yolov8, README.md:200-207).
# Variant names here: '11n', '11s', '11m', '11l', '11x'.  The layer table below reproduces the published parameter
# counts of all five scales exactly (tests/test_weights_cpu.py::test_yolo11_parameter_counts).
SCALES11 = {'n': (0.50, 0.25, 1024), 's': (0.50, 0.50, 1024), 'm': (0.50, 1.00, 512), 'l': (1.00, 1.00, 512),
            'x': (1.00, 1.50, 512)}


def is_yolo11(variant):
    return str(variant).startswith('11')


def arch11(variant):
    """Channel widths / repeats / block kinds of yolo11{n,s,m,l,x}.  c3k: C3k2 blocks of layers 2, 4, 13, 16, 19 use
    C3k inner blocks only for the m/l/x scales (parse_model forces c3k=True there); layers 6, 8, 22 always do."""
    depth, width, maxc = SCALES11[variant[2:]]
    ch = lambda c: make_divisible(min(c, maxc) * width, 8)
    return dict(c64=ch(64), c128=ch(128), c256=ch(256), c512=ch(512), c1024=ch(1024), n=max(round(2 * depth), 1),
                c3k=variant[2:] in 'mlx')


def conv_layers11(variant, nc=5):
    """Ordered list of (prefix, cin, cout, k, groups, has_bn) for every conv of yolo11 in forward order (Conv modules
    have BN + SiLU unless the block says otherwise; the two head outputs per level are plain Conv2d with bias)."""
    a = arch11(variant)
    c64, c128, c256, c512, c1024, n, big = a['c64'], a['c128'], a['c256'], a['c512'], a['c1024'], a['n'], a['c3k']
    L = []

    def conv(p, cin, cout, k, g=1, bn=True):
        L.append((p, cin, cout, k, g, bn))

    def c3k2(p, cin, cout, c3k, e=0.5):
        c = int(cout * e)
        conv(p + '.cv1', cin, 2 * c, 1)
        for i in range(n):
            m = '%s.m.%d' % (p, i)
            if c3k:     # C3k(c, c, 2): cv1, cv2 -> c/2; two Bottleneck(c/2, c/2, e=1.0); cv3
                c_ = c // 2
                conv(m + '.cv1', c, c_, 1)
                conv(m + '.cv2', c, c_, 1)
                for j in range(2):
                    conv('%s.m.%d.cv1' % (m, j), c_, c_, 3)
                    conv('%s.m.%d.cv2' % (m, j), c_, c_, 3)
                conv(m + '.cv3', 2 * c_, c, 1)
            else:       # Bottleneck(c, c, e=0.5)
                conv(m + '.cv1', c, c // 2, 3)
                conv(m + '.cv2', c // 2, c, 3)
        conv(p + '.cv2', (2 + n) * c, cout, 1)

    conv('model.0', 3, c64, 3)
    conv('model.1', c64, c128, 3)
    c3k2('model.2', c128, c256, big, 0.25)
    conv('model.3', c256, c256, 3)
    c3k2('model.4', c256, c512, big, 0.25)
    conv('model.5', c512, c512, 3)
    c3k2('model.6', c512, c512, True)
    conv('model.7', c512, c1024, 3)
    c3k2('model.8', c1024, c1024, True)
    conv('model.9.cv1', c1024, c1024 // 2, 1)
    conv('model.9.cv2', c1024 * 2, c1024, 1)
    c = c1024 // 2                                        # C2PSA(c1024, c1024, n, e=0.5)
    conv('model.10.cv1', c1024, 2 * c, 1)
    for i in range(n):
        m = 'model.10.m.%d' % i
        nh = c // 64
        conv(m + '.attn.qkv', c, c + 2 * nh * 32, 1)      # head_dim 64, key_dim 32 (attn_ratio 0.5); no activation
        conv(m + '.attn.pe', c, c, 3, g=c)                # depthwise positional conv, no activation
        conv(m + '.attn.proj', c, c, 1)                   # no activation
        conv(m + '.ffn.0', c, 2 * c, 1)
        conv(m + '.ffn.1', 2 * c, c, 1)                   # no activation
    conv('model.10.cv2', 2 * c, c1024, 1)
    c3k2('model.13', c1024 + c512, c512, big)
    c3k2('model.16', c512 + c512, c256, big)
    conv('model.17', c256, c256, 3)
    c3k2('model.19', c256 + c512, c512, big)
    conv('model.20', c512, c512, 3)
    c3k2('model.22', c512 + c1024, c1024, True)
    cb = max(16, c256 // 4, 64)
    cc = max(c256, min(nc, 100))
    for l, cl in enumerate((c256, c512, c1024)):
        conv('model.23.cv2.%d.0' % l, cl, cb, 3)
        conv('model.23.cv2.%d.1' % l, cb, cb, 3)
        conv('model.23.cv2.%d.2' % l, cb, 64, 1, bn=False)
        conv('model.23.cv3.%d.0.0' % l, cl, cl, 3, g=cl)  # DWConv
        conv('model.23.cv3.%d.0.1' % l, cl, cc, 1)
        conv('model.23.cv3.%d.1.0' % l, cc, cc, 3, g=cc)  # DWConv
        conv('model.23.cv3.%d.1.1' % l, cc, cc, 1)
        conv('model.23.cv3.%d.2' % l, cc, nc, 1, bn=False)
    return L


def count_parameters11(variant, nc=80):
    """nn.Module parameter count of the unfused ultralytics model (conv weights, BN weight + bias, head biases, DFL)."""
    tot = 16
    for (p, cin, cout, k, g, bn) in conv_layers11(variant, nc):
        tot += cout * (cin // g) * k * k + (2 * cout if bn else cout)
    return tot


def make_random_weights11(variant='11n', nc=5, seed=0, cls_bias=-3.0, calibration='auto'):
    """Seeded random-init yolo11 with the ultralytics state-dict key names (same recipe as make_random_weights: BN
    running statistics from init_calibration.json for the (variant, seed) pairs tests/diag/calibrate_init.py lists)."""
    calib = _load_calibration(variant, seed) if calibration == 'auto' else calibration
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for (p, cin, cout, k, grp, bn) in conv_layers11(variant, nc):
        fan_in = (cin // grp) * k * k
        if bn:
            sd[p + '.conv.weight'] = torch.randn(cout, cin // grp, k, k, generator=g) / math.sqrt(fan_in)
            sd[p + '.bn.weight'] = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sd[p + '.bn.bias'] = 0.1 * torch.randn(cout, generator=g)
            mu, var = (calib[p] if calib and p in calib else (0.0, 1.0))
            sd[p + '.bn.running_mean'] = torch.full((cout,), float(mu))
            sd[p + '.bn.running_var'] = float(var) * (0.8 + 0.4 * torch.rand(cout, generator=g))
        elif '.cv2.' in p:
            sd[p + '.weight'] = torch.randn(cout, cin, 1, 1, generator=g) * (1.5 / math.sqrt(cin))
            sd[p + '.bias'] = (-0.6 * torch.arange(16, dtype=torch.float32)).repeat(4) + 1.0
        else:
            sd[p + '.weight'] = torch.randn(cout, cin, 1, 1, generator=g) * (2.5 / math.sqrt(cin))
            sd[p + '.bias'] = torch.full((cout,), float(cls_bias))
    sd['model.23.dfl.conv.weight'] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return {'format': FORMAT, 'variant': variant, 'nc': nc, 'names': dict(CLASS_NAMES) if nc == 5 else
            {i: 'class%d' % i for i in range(nc)}, 'state_dict': sd}


def _load_calibration(variant, seed):
    if os.path.exists(_CALIB_PATH):
        with open(_CALIB_PATH) as f:
            t = json.load(f)
        return t.get('%s:%d' % (variant, seed))
    return None


# Random-init recipes.  Both draw the same backbone (conv weights ~ N(0, 1/fan_in), calibrated BN); they
# differ in the Detect head only.
#   'v1' (round 1): random DFL head (boxes of random aspect, 2-10 cells), all three pyramid levels fire, class gain 2.5
#        (sigmoid saturates: scores 0.99-1.0).  Overlapping candidates with near-tied scores everywhere -- NMS and the
#        IoU-graph merge pick the winner of a near-tie, which any change of arithmetic flips.
#   'v2' ("separated peaks"): a well-conditioned decision structure for the end-to-end acceptance test --
#        * the DFL head decodes every anchor to (almost) the same square box of 2 x 2.1 cells: neighbouring candidates
#          overlap by FIXED IoUs, (1,0) 0.62, (1,1) 0.42, (2,0) 0.36, (2,1) 0.26, all >= 14 % away from the NMS (0.5)
#          and soft-merge (0.3) thresholds, so a blob of candidates collapses to its single best member;
#        * class gain 1.0: logits of the detections spread over 0..7, where the fp32 sigmoid still resolves them;
#        * only the stride-8 level fires (class bias -100 on P4 / P5): no cross-level pairs sitting at IoU 0.25.
#        What is left is the irreducible part: score-threshold crossings and true near-ties of the two best members of
#        a blob, ~0.2 % of the sources with fp16 storage (tests/diag/recipe_probe.py).
RECIPES = {
    'v1': dict(bn_beta=0.0, bn_beta_jit=0.1, dfl_gain=1.5, dfl_center=None, dfl_beta=0.6, cls_gain=2.5, levels=None),
    'v2': dict(bn_beta=0.0, bn_beta_jit=0.1, dfl_gain=0.05, dfl_center=2.14, dfl_beta=2.0, cls_gain=1.0, levels=(0,)),
}
# class-branch bias of the v2 recipe per variant: ~8 detections per 512^2 tile of the synthetic mosaics at scoreThr 0.5
V2_CLS_BIAS = {'n': -4.6, 'l': -4.6}


def _load_calibration_key(key):
    if os.path.exists(_CALIB_PATH):
        with open(_CALIB_PATH) as f:
            return json.load(f).get(key)
    return None


def calibration_key(variant, seed, recipe='v1'):
    """The BN calibration depends on the backbone draw only, which the recipes share."""
    return '%s:%d' % (variant, seed)


def make_random_weights(variant='n', nc=5, seed=0, cls_bias=-3.0, calibration='auto', recipe='v1', **knobs):
    """Seeded random-init YOLOv8 (our own init; the reference has none).  Conv weights ~ N(0, 1/fan_in); the BN
    running statistics come from init_calibration.json (per-layer scalar mean/var of the pre-BN activation on a
    synthetic preprocessed tile, produced by tests/diag/calibrate_init.py for the (recipe, variant, seed) triples it
    lists; others fall back to mean 0 / var 1) so that every layer works at unit scale and detections depend on the
    image.  The class-branch bias sets the candidate density; the DFL bias sets the box size (RECIPES above)."""
    if is_yolo11(variant):
        return make_random_weights11(variant, nc, seed, cls_bias, calibration)
    r = dict(RECIPES[recipe])
    r.update(knobs)
    calib = _load_calibration_key(calibration_key(variant, seed, recipe)) if calibration == 'auto' else calibration
    g = torch.Generator().manual_seed(seed)
    layers, cb, cc = conv_bn_layers(variant, nc)
    sd = {}
    for (p, cin, cout, k) in layers:
        fan_in = cin * k * k
        sd[p + '.conv.weight'] = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(fan_in)
        mu, var = (calib[p] if calib and p in calib else (0.0, 1.0))
        sd[p + '.bn.weight'] = 1.0 + 0.1 * torch.randn(cout, generator=g)
        sd[p + '.bn.bias'] = r['bn_beta'] + r['bn_beta_jit'] * torch.randn(cout, generator=g)
        sd[p + '.bn.running_mean'] = torch.full((cout,), float(mu))
        sd[p + '.bn.running_var'] = float(var) * (0.8 + 0.4 * torch.rand(cout, generator=g))
    k16 = torch.arange(16, dtype=torch.float32)
    if r['dfl_center'] is None:
        dfl_bias = (-r['dfl_beta'] * k16).repeat(4) + 1.0
    else:
        dfl_bias = (-r['dfl_beta'] * (k16 - float(r['dfl_center'])) ** 2).repeat(4)
    for l in range(3):
        sd['model.22.cv2.%d.2.weight' % l] = torch.randn(64, cb, 1, 1, generator=g) * (r['dfl_gain'] / math.sqrt(cb))
        sd['model.22.cv2.%d.2.bias' % l] = dfl_bias.clone()
        sd['model.22.cv3.%d.2.weight' % l] = torch.randn(nc, cc, 1, 1, generator=g) * (r['cls_gain'] / math.sqrt(cc))
        on = r.get('levels') is None or l in r['levels']
        sd['model.22.cv3.%d.2.bias' % l] = torch.full((nc,), float(cls_bias) if on else -100.0)
    sd['model.22.dfl.conv.weight'] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return {'format': FORMAT, 'variant': variant, 'nc': nc, 'names': dict(CLASS_NAMES), 'state_dict': sd}


def save_weights(weights, path):
    torch.save(weights, path)


class _StubPickle(object):
    """`pickle_module` for torch.load: classes / functions whose module cannot be imported here (ultralytics.*, and
    whatever a training environment left inside the checkpoint) become stand-in torch.nn.Module subclasses, so the
    module tree — `_modules`, `_parameters`, `_buffers` live in the pickled `__dict__` — is rebuilt without their code."""
    __name__ = 'caesar_yolo_b200._stub_pickle'
    _cache = {}

    @classmethod
    def _stub(cls, module, name):
        key = (module, name)
        if key not in cls._cache:
            def __init__(self, *args, **kwargs):
                torch.nn.Module.__init__(self)
            cls._cache[key] = type(name, (torch.nn.Module,), {'__module__': module, '__init__': __init__,
                                                             '_caesar_stub': True})
        return cls._cache[key]

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                return _StubPickle._stub(module, name)

    load = staticmethod(pickle.load)
    dump = staticmethod(pickle.dump)
    Pickler = pickle.Pickler


_STEM_WIDTH_TO_VARIANT = {16: 'n', 32: 's', 48: 'm', 64: 'l', 80: 'x'}


def weights_from_state_dict(sd, names=None):
    """ultralytics-keyed YOLOv8 detection state dict -> this build's weight dict.  Infers the variant from the stem
    width and nc from the class branch, checks every tensor the device model needs (name and shape) and rewrites
    fused Conv modules (conv.weight + conv.bias, no bn.*) as Conv + identity BN."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items() if torch.is_tensor(v)}
    w0 = sd.get('model.0.conv.weight')
    if w0 is None or w0.dim() != 4 or tuple(w0.shape[1:]) != (3, 3, 3) or int(w0.shape[0]) not in _STEM_WIDTH_TO_VARIANT:
        raise ValueError("not a YOLOv8 detection checkpoint: model.0.conv.weight should be [16|32|48|64|80, 3, 3, 3]")
    if 'model.23.cv3.0.2.weight' in sd and 'model.10.m.0.attn.qkv.conv.weight' in sd:
        return _weights11_from_state_dict(sd, names)
    variant = _STEM_WIDTH_TO_VARIANT[int(w0.shape[0])]
    head = sd.get('model.22.cv3.0.2.weight')
    if head is None:
        raise ValueError("not a YOLOv8 / YOLO11 detection checkpoint: no Detect head at model.22 / model.23 "
                         "(segmentation / pose / other families are not supported by this build)")
    nc = int(head.shape[0])
    layers, cb, cc = conv_bn_layers(variant, nc)
    out = {}

    def need(key, shape):
        t = sd.get(key)
        if t is None:
            raise ValueError("checkpoint does not match yolov8%s (nc=%d): missing tensor %s" % (variant, nc, key))
        if tuple(t.shape) != tuple(shape):
            raise ValueError("checkpoint does not match yolov8%s (nc=%d): %s has shape %s, expected %s"
                             % (variant, nc, key, tuple(t.shape), tuple(shape)))
        return t
    for (p, cin, cout, k) in layers:
        out[p + '.conv.weight'] = need(p + '.conv.weight', (cout, cin, k, k))
        if p + '.bn.weight' in sd:
            for f in ('weight', 'bias', 'running_mean', 'running_var'):
                out['%s.bn.%s' % (p, f)] = need('%s.bn.%s' % (p, f), (cout,))
        else:   # fused: y = conv(x) + bias.  Identity BN under eps = 1e-3: gamma / sqrt(var + eps) = 1
            out[p + '.bn.weight'] = torch.ones(cout)
            out[p + '.bn.bias'] = need(p + '.conv.bias', (cout,))
            out[p + '.bn.running_mean'] = torch.zeros(cout)
            out[p + '.bn.running_var'] = torch.full((cout,), 1.0 - 1e-3)
    for l in range(3):
        out['model.22.cv2.%d.2.weight' % l] = need('model.22.cv2.%d.2.weight' % l, (64, cb, 1, 1))
        out['model.22.cv2.%d.2.bias' % l] = need('model.22.cv2.%d.2.bias' % l, (64,))
        out['model.22.cv3.%d.2.weight' % l] = need('model.22.cv3.%d.2.weight' % l, (nc, cc, 1, 1))
        out['model.22.cv3.%d.2.bias' % l] = need('model.22.cv3.%d.2.bias' % l, (nc,))
    dfl = sd.get('model.22.dfl.conv.weight')
    if dfl is not None and not torch.equal(dfl.flatten(), torch.arange(16, dtype=torch.float32)):
        raise ValueError("model.22.dfl.conv.weight is not arange(16): unsupported DFL")
    out['model.22.dfl.conv.weight'] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    if names is None:
        names = dict(CLASS_NAMES) if nc == len(CLASS_NAMES) else {i: 'class%d' % i for i in range(nc)}
    elif not isinstance(names, dict):
        names = {i: n for i, n in enumerate(names)}
    names = {int(k): str(v) for k, v in names.items()}
    if sorted(names) != list(range(nc)):
        raise ValueError("checkpoint names %r do not cover nc=%d classes" % (names, nc))
    return {'format': FORMAT, 'variant': variant, 'nc': nc, 'names': names, 'state_dict': out}


def _names_dict(names, nc):
    if names is None:
        names = dict(CLASS_NAMES) if nc == len(CLASS_NAMES) else {i: 'class%d' % i for i in range(nc)}
    elif not isinstance(names, dict):
        names = {i: n for i, n in enumerate(names)}
    names = {int(k): str(v) for k, v in names.items()}
    if sorted(names) != list(range(nc)):
        raise ValueError("checkpoint names %r do not cover nc=%d classes" % (names, nc))
    return names


def _weights11_from_state_dict(sd, names):
    """yolo11 detection state dict -> weight dict (variant from the stem width + the depth of model.2 / block kind)."""
    c0 = int(sd['model.0.conv.weight'].shape[0])
    nc = int(sd['model.23.cv3.0.2.weight'].shape[0])
    cands = [v for v in ('11n', '11s', '11m', '11l', '11x') if arch11(v)['c64'] == c0]
    # m and l share widths (64): l has n = 2 repeats
    variant = None
    for v in cands:
        if all((p + ('.conv.weight' if bn else '.weight')) in sd for (p, _, _, _, _, bn) in conv_layers11(v, nc)) and \
                ('model.2.m.%d.cv1.conv.weight' % arch11(v)['n']) not in sd:
            variant = v
            break
    if variant is None:
        raise ValueError("checkpoint does not match any yolo11 scale (stem width %d)" % c0)
    out = {}
    for (p, cin, cout, k, g, bn) in conv_layers11(variant, nc):
        if bn:
            w = sd[p + '.conv.weight']
            if tuple(w.shape) != (cout, cin // g, k, k):
                raise ValueError("checkpoint does not match yolo%s (nc=%d): %s.conv.weight has shape %s, expected %s"
                                 % (variant, nc, p, tuple(w.shape), (cout, cin // g, k, k)))
            out[p + '.conv.weight'] = w
            if p + '.bn.weight' in sd:
                for f in ('weight', 'bias', 'running_mean', 'running_var'):
                    out['%s.bn.%s' % (p, f)] = sd['%s.bn.%s' % (p, f)]
            else:
                out[p + '.bn.weight'] = torch.ones(cout)
                out[p + '.bn.bias'] = sd[p + '.conv.bias']
                out[p + '.bn.running_mean'] = torch.zeros(cout)
                out[p + '.bn.running_var'] = torch.full((cout,), 1.0 - 1e-3)
        else:
            out[p + '.weight'] = sd[p + '.weight']
            out[p + '.bias'] = sd[p + '.bias']
    out['model.23.dfl.conv.weight'] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return {'format': FORMAT, 'variant': variant, 'nc': nc, 'names': _names_dict(names, nc), 'state_dict': out}


def load_ultralytics_checkpoint(path):
    """Reads an ultralytics YOLOv8 detection `.pt` without the ultralytics package (see _StubPickle).  Follows
    ultralytics' own loader: the EMA model when present, else 'model'; weights converted to fp32 (they are stored as
    fp16).  Also accepts a bare module or a bare state dict."""
    ck = torch.load(path, map_location='cpu', weights_only=False, pickle_module=_StubPickle)
    if isinstance(ck, dict) and ck.get('format') == FORMAT:
        return ck
    obj = ck
    if isinstance(ck, dict) and ('model' in ck or 'ema' in ck):
        obj = ck.get('ema') if ck.get('ema') is not None else ck.get('model')
        if obj is None:
            obj = ck
    if isinstance(obj, torch.nn.Module):
        names = getattr(obj, 'names', None)
        return weights_from_state_dict(obj.state_dict(), names)
    if isinstance(obj, dict):
        sd = obj.get('state_dict', obj)
        return weights_from_state_dict(sd, obj.get('names') if isinstance(obj.get('names'), (dict, list)) else None)
    raise ValueError("%s: unsupported checkpoint layout (%s)" % (path, type(obj).__name__))


def load_weights(path):
    """`--weights`: this build's weight file or an ultralytics YOLOv8 detection checkpoint."""
    return load_ultralytics_checkpoint(path)
