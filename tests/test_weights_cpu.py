"""`--weights` loading (scripts/run.py:347 `YOLO(weights_path)`): ultralytics YOLOv8 `.pt` checkpoints are read without
the ultralytics package.  The checkpoint is produced here by pickling a module tree whose classes live in a throw-away
`ultralytics.*` package (same attribute layout as ultralytics' Conv / C2f / Bottleneck / SPPF / Concat / Detect / DFL /
DetectionModel), which is removed from sys.modules before loading — the situation on a box without ultralytics."""
import sys
import types

import pytest
import torch
import torch.nn as nn

from caesar_yolo_b200 import weights as W

_MODS = ['ultralytics', 'ultralytics.nn', 'ultralytics.nn.tasks', 'ultralytics.nn.modules',
         'ultralytics.nn.modules.conv', 'ultralytics.nn.modules.block', 'ultralytics.nn.modules.head',
         'ultralytics.utils']


def _fake_package():
    mods = {n: types.ModuleType(n) for n in _MODS}

    def reg(modname):
        def deco(cls):
            cls.__module__ = modname
            cls.__qualname__ = cls.__name__
            setattr(mods[modname], cls.__name__, cls)
            return cls
        return deco

    @reg('ultralytics.nn.modules.conv')
    class Conv(nn.Module):
        def __init__(self, c1, c2, k=1, s=1):
            super().__init__()
            self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
            self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
            self.act = nn.SiLU()

    @reg('ultralytics.nn.modules.conv')
    class Concat(nn.Module):
        def __init__(self, d=1):
            super().__init__()
            self.d = d

    @reg('ultralytics.nn.modules.block')
    class Bottleneck(nn.Module):
        def __init__(self, c, shortcut):
            super().__init__()
            self.cv1, self.cv2, self.add = Conv(c, c, 3), Conv(c, c, 3), shortcut

    @reg('ultralytics.nn.modules.block')
    class C2f(nn.Module):
        def __init__(self, c1, c2, n, shortcut):
            super().__init__()
            self.c = c2 // 2
            self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1), Conv((2 + n) * self.c, c2, 1)
            self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    @reg('ultralytics.nn.modules.block')
    class SPPF(nn.Module):
        def __init__(self, c1, c2):
            super().__init__()
            self.cv1, self.cv2 = Conv(c1, c1 // 2, 1), Conv(c1 * 2, c2, 1)
            self.m = nn.MaxPool2d(5, 1, 2)

    @reg('ultralytics.nn.modules.block')
    class DFL(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(16, 1, 1, bias=False).requires_grad_(False)
            self.conv.weight.data[:] = torch.arange(16, dtype=torch.float).view(1, 16, 1, 1)

    @reg('ultralytics.nn.modules.head')
    class Detect(nn.Module):
        def __init__(self, nc, ch, cb, cc):
            super().__init__()
            self.nc, self.nl, self.reg_max = nc, 3, 16
            self.stride = torch.tensor([8., 16., 32.])
            self.cv2 = nn.ModuleList(nn.Sequential(Conv(c, cb, 3), Conv(cb, cb, 3), nn.Conv2d(cb, 64, 1)) for c in ch)
            self.cv3 = nn.ModuleList(nn.Sequential(Conv(c, cc, 3), Conv(cc, cc, 3), nn.Conv2d(cc, nc, 1)) for c in ch)
            self.dfl = DFL()

    @reg('ultralytics.utils')
    class IterableSimpleNamespace(object):   # a non-module object ultralytics leaves in checkpoints
        def __init__(self, **kw):
            self.__dict__.update(kw)

    @reg('ultralytics.nn.tasks')
    class DetectionModel(nn.Module):
        def __init__(self, variant, nc, names):
            super().__init__()
            a = W.arch(variant)
            c1, c2, c3, c4, c5 = a['c1'], a['c2'], a['c3'], a['c4'], a['c5']
            cb, cc = max(16, c3 // 4, 64), max(c3, min(nc, 100))
            up = lambda: nn.Upsample(None, 2, 'nearest')
            self.model = nn.Sequential(
                Conv(3, c1, 3, 2), Conv(c1, c2, 3, 2), C2f(c2, c2, a['n2'], True), Conv(c2, c3, 3, 2),
                C2f(c3, c3, a['n4'], True), Conv(c3, c4, 3, 2), C2f(c4, c4, a['n6'], True), Conv(c4, c5, 3, 2),
                C2f(c5, c5, a['n8'], True), SPPF(c5, c5), up(), Concat(), C2f(c5 + c4, c4, a['nh'], False), up(),
                Concat(), C2f(c4 + c3, c3, a['nh'], False), Conv(c3, c3, 3, 2), Concat(),
                C2f(c3 + c4, c4, a['nh'], False), Conv(c4, c4, 3, 2), Concat(), C2f(c4 + c5, c5, a['nh'], False),
                Detect(nc, (c3, c4, c5), cb, cc))
            self.names = names
            self.nc = nc
            self.args = IterableSimpleNamespace(imgsz=640, task='detect')
            self.yaml = {'nc': nc, 'scale': variant}
    return mods, DetectionModel


def _save_checkpoint(path, variant, nc, names, seed, half=True, ema=False):
    mods, DetectionModel = _fake_package()
    ours = W.make_random_weights(variant, nc, seed=seed)
    sys.modules.update(mods)
    try:
        m = DetectionModel(variant, nc, names)
        missing, unexpected = m.load_state_dict(ours['state_dict'], strict=False)
        assert not unexpected and all(k.endswith('num_batches_tracked') for k in missing), (missing, unexpected)
        if half:
            m = m.half()
        ck = {'epoch': -1, 'best_fitness': None, 'model': None if ema else m, 'ema': m if ema else None,
              'updates': 7, 'optimizer': None, 'train_args': {'model': 'yolov8%s.yaml' % variant, 'imgsz': 640},
              'date': '2024-01-01', 'version': '8.3.0'}
        torch.save(ck, path)
    finally:
        for n in _MODS:
            sys.modules.pop(n, None)
    return ours


@pytest.mark.parametrize("variant,nc,ema", [('n', 5, False), ('s', 5, True), ('l', 3, False)])
def test_ultralytics_checkpoint_loads_without_ultralytics(tmp_path, variant, nc, ema):
    names = {i: 'c%d' % i for i in range(nc)} if nc != 5 else list(W.CLASS_NAMES.values())
    p = str(tmp_path / 'yolov8.pt')
    ours = _save_checkpoint(p, variant, nc, names, seed=3, ema=ema)
    assert 'ultralytics' not in sys.modules
    w = W.load_weights(p)
    assert w['format'] == W.FORMAT and w['variant'] == variant and w['nc'] == nc
    assert w['names'] == ({i: n for i, n in enumerate(names)} if isinstance(names, list) else names)
    sd = w['state_dict']
    assert set(sd) == set(ours['state_dict'])
    for k, v in ours['state_dict'].items():
        assert sd[k].dtype == torch.float32
        assert torch.equal(sd[k], v.half().float()), k     # checkpoints store fp16


def test_fused_checkpoint_is_conv_plus_identity_bn(tmp_path):
    """`model.fuse()`d checkpoints carry conv.weight + conv.bias and no BatchNorm: they load as Conv + identity BN."""
    ours = W.make_random_weights('n', 5, seed=1)
    fused = {}
    for k, v in ours['state_dict'].items():
        if '.bn.' in k:
            continue
        fused[k] = v
        if k.endswith('.conv.weight') and not k.startswith('model.22.dfl'):
            p = k[:-len('.conv.weight')]
            g, b = ours['state_dict'][p + '.bn.weight'], ours['state_dict'][p + '.bn.bias']
            mu, var = ours['state_dict'][p + '.bn.running_mean'], ours['state_dict'][p + '.bn.running_var']
            s = g / torch.sqrt(var + 1e-3)
            fused[k] = v * s.view(-1, 1, 1, 1)
            fused[p + '.conv.bias'] = b - mu * s
    path = str(tmp_path / 'fused.pt')
    torch.save(fused, path)
    w = W.load_weights(path)
    sd = w['state_dict']
    for k in fused:
        if k.endswith('.conv.bias'):
            p = k[:-len('.conv.bias')]
            s = sd[p + '.bn.weight'] / torch.sqrt(sd[p + '.bn.running_var'] + 1e-3)
            assert torch.allclose(s, torch.ones_like(s), atol=1e-6)
            assert torch.allclose(sd[p + '.bn.bias'] - sd[p + '.bn.running_mean'] * s, fused[k], atol=1e-6)
            assert torch.equal(sd[p + '.conv.weight'], fused[p + '.conv.weight'])


def test_foreign_checkpoints_are_rejected_with_a_reason(tmp_path):
    ours = W.make_random_weights('n', 5, seed=0)['state_dict']
    p = str(tmp_path / 'x.pt')
    # yolo11-style: no Detect at model.22
    sd = {k.replace('model.22.', 'model.23.'): v for k, v in ours.items()}
    torch.save(sd, p)
    with pytest.raises(ValueError, match="no Detect head at model.22"):
        W.load_weights(p)
    # a layer of the wrong depth (yolov8s stem width with yolov8n body)
    sd = dict(ours)
    sd['model.0.conv.weight'] = torch.zeros(32, 3, 3, 3)
    torch.save(sd, p)
    with pytest.raises(ValueError, match="does not match yolov8s"):
        W.load_weights(p)
    # missing tensor
    sd = dict(ours)
    del sd['model.4.m.1.cv2.conv.weight']
    torch.save(sd, p)
    with pytest.raises(ValueError, match="missing tensor model.4.m.1.cv2.conv.weight"):
        W.load_weights(p)
    torch.save({'hello': 1}, p)
    with pytest.raises(ValueError):
        W.load_weights(p)


def test_own_weight_file_roundtrip(tmp_path):
    w = W.make_random_weights('n', 5, seed=2)
    p = str(tmp_path / 'w.pt')
    W.save_weights(w, p)
    w2 = W.load_weights(p)
    assert w2['variant'] == 'n' and w2['names'] == W.CLASS_NAMES
    assert all(torch.equal(w2['state_dict'][k], v) for k, v in w['state_dict'].items())


def test_yolo11_parameter_counts():
    """The yolo11 layer table (shared by the device model's graph builder and the oracle) reproduces the parameter
    counts ultralytics publishes for the five scales at nc = 80 (model summaries of yolo11{n,s,m,l,x}.pt)."""
    want = {'11n': 2624080, '11s': 9458752, '11m': 20114688, '11l': 25372160, '11x': 56966176}
    for v, n in want.items():
        assert W.count_parameters11(v, 80) == n, v


def test_yolo11_state_dict_roundtrip_and_variant_inference(tmp_path):
    for v in ('11n', '11m', '11l'):
        w = W.make_random_weights(v, 5, seed=0)
        p = str(tmp_path / ('%s.pt' % v))
        torch.save({'model': None, 'ema': None, 'state_dict': w['state_dict'], 'names': list(W.CLASS_NAMES.values())}, p)
        w2 = W.load_weights(p)
        assert w2['variant'] == v and w2['nc'] == 5 and w2['names'] == W.CLASS_NAMES
        assert set(w2['state_dict']) == set(w['state_dict'])
    sd = dict(W.make_random_weights('11n', 5, seed=0)['state_dict'])
    del sd['model.10.m.0.ffn.1.conv.weight']
    torch.save(sd, str(tmp_path / 'bad.pt'))
    with pytest.raises(ValueError, match="does not match any yolo11 scale"):
        W.load_weights(str(tmp_path / 'bad.pt'))


def _fake_package11():
    """Throw-away `ultralytics.*` modules with the attribute layout of the YOLO11 blocks (Conv / DWConv / Bottleneck /
    C3k / C3k2 / SPPF / Attention / PSABlock / C2PSA / Detect with the depthwise-separable class branch)."""
    mods = {n: types.ModuleType(n) for n in _MODS}

    def reg(modname):
        def deco(cls):
            cls.__module__ = modname
            cls.__qualname__ = cls.__name__
            setattr(mods[modname], cls.__name__, cls)
            return cls
        return deco

    @reg('ultralytics.nn.modules.conv')
    class Conv(nn.Module):
        def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
            super().__init__()
            self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=False)
            self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
            self.act = nn.SiLU() if act else nn.Identity()

    @reg('ultralytics.nn.modules.conv')
    class DWConv(Conv):
        def __init__(self, c1, c2, k=1, s=1):
            super().__init__(c1, c2, k, s, g=c1)

    @reg('ultralytics.nn.modules.conv')
    class Concat(nn.Module):
        def __init__(self, d=1):
            super().__init__()
            self.d = d

    @reg('ultralytics.nn.modules.block')
    class Bottleneck(nn.Module):
        def __init__(self, c1, c2, e=0.5):
            super().__init__()
            c_ = int(c2 * e)
            self.cv1, self.cv2, self.add = Conv(c1, c_, 3), Conv(c_, c2, 3), True

    @reg('ultralytics.nn.modules.block')
    class C3k(nn.Module):
        def __init__(self, c1, c2, n=2):
            super().__init__()
            c_ = int(c2 * 0.5)
            self.cv1, self.cv2, self.cv3 = Conv(c1, c_, 1), Conv(c1, c_, 1), Conv(2 * c_, c2, 1)
            self.m = nn.Sequential(*(Bottleneck(c_, c_, e=1.0) for _ in range(n)))

    @reg('ultralytics.nn.modules.block')
    class C3k2(nn.Module):
        def __init__(self, c1, c2, n, c3k, e=0.5):
            super().__init__()
            self.c = int(c2 * e)
            self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1), Conv((2 + n) * self.c, c2, 1)
            self.m = nn.ModuleList(C3k(self.c, self.c, 2) if c3k else Bottleneck(self.c, self.c) for _ in range(n))

    @reg('ultralytics.nn.modules.block')
    class SPPF(nn.Module):
        def __init__(self, c1, c2):
            super().__init__()
            self.cv1, self.cv2 = Conv(c1, c1 // 2, 1), Conv(c1 * 2, c2, 1)
            self.m = nn.MaxPool2d(5, 1, 2)

    @reg('ultralytics.nn.modules.block')
    class Attention(nn.Module):
        def __init__(self, dim, num_heads, attn_ratio=0.5):
            super().__init__()
            self.num_heads, self.head_dim = num_heads, dim // num_heads
            self.key_dim = int(self.head_dim * attn_ratio)
            self.qkv = Conv(dim, dim + 2 * self.key_dim * num_heads, 1, act=False)
            self.proj = Conv(dim, dim, 1, act=False)
            self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    @reg('ultralytics.nn.modules.block')
    class PSABlock(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.attn = Attention(c, c // 64)
            self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))

    @reg('ultralytics.nn.modules.block')
    class C2PSA(nn.Module):
        def __init__(self, c1, n):
            super().__init__()
            self.c = c1 // 2
            self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1), Conv(2 * self.c, c1, 1)
            self.m = nn.Sequential(*(PSABlock(self.c) for _ in range(n)))

    @reg('ultralytics.nn.modules.block')
    class DFL(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(16, 1, 1, bias=False).requires_grad_(False)
            self.conv.weight.data[:] = torch.arange(16, dtype=torch.float).view(1, 16, 1, 1)

    @reg('ultralytics.nn.modules.head')
    class Detect(nn.Module):
        def __init__(self, nc, ch, cb, cc):
            super().__init__()
            self.nc = nc
            self.cv2 = nn.ModuleList(nn.Sequential(Conv(c, cb, 3), Conv(cb, cb, 3), nn.Conv2d(cb, 64, 1)) for c in ch)
            self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(DWConv(c, c, 3), Conv(c, cc, 1)),
                                                   nn.Sequential(DWConv(cc, cc, 3), Conv(cc, cc, 1)),
                                                   nn.Conv2d(cc, nc, 1)) for c in ch)
            self.dfl = DFL()

    @reg('ultralytics.nn.tasks')
    class DetectionModel(nn.Module):
        def __init__(self, variant, nc, names):
            super().__init__()
            a = W.arch11(variant)
            c64, c128, c256, c512, c1024, n, big = (a['c64'], a['c128'], a['c256'], a['c512'], a['c1024'], a['n'],
                                                    a['c3k'])
            cb, cc = max(16, c256 // 4, 64), max(c256, min(nc, 100))
            up = lambda: nn.Upsample(None, 2, 'nearest')
            self.model = nn.Sequential(
                Conv(3, c64, 3, 2), Conv(c64, c128, 3, 2), C3k2(c128, c256, n, big, 0.25), Conv(c256, c256, 3, 2),
                C3k2(c256, c512, n, big, 0.25), Conv(c512, c512, 3, 2), C3k2(c512, c512, n, True),
                Conv(c512, c1024, 3, 2), C3k2(c1024, c1024, n, True), SPPF(c1024, c1024), C2PSA(c1024, n), up(),
                Concat(), C3k2(c1024 + c512, c512, n, big), up(), Concat(), C3k2(c512 + c512, c256, n, big),
                Conv(c256, c256, 3, 2), Concat(), C3k2(c256 + c512, c512, n, big), Conv(c512, c512, 3, 2), Concat(),
                C3k2(c512 + c1024, c1024, n, True), Detect(nc, (c256, c512, c1024), cb, cc))
            self.names = names
    return mods, DetectionModel


@pytest.mark.parametrize("variant", ['11n', '11l'])
def test_yolo11_checkpoint_loads_without_ultralytics(tmp_path, variant):
    """A pickled YOLO11 DetectionModel (module tree with the ultralytics attribute names, built from throw-away
    classes) loads through the stand-in unpickler; the state-dict keys of that tree are exactly the keys of this build's
    layer table, and the parameter count of the tree at nc = 80 is the published one."""
    mods, DetectionModel = _fake_package11()
    ours = W.make_random_weights(variant, 5, seed=4)
    p = str(tmp_path / 'yolo11.pt')
    sys.modules.update(mods)
    try:
        m = DetectionModel(variant, 5, dict(W.CLASS_NAMES))
        missing, unexpected = m.load_state_dict(ours['state_dict'], strict=False)
        assert not unexpected and all(k.endswith('num_batches_tracked') for k in missing), (missing, unexpected)
        m80 = DetectionModel(variant, 80, {i: str(i) for i in range(80)})
        assert sum(p_.numel() for p_ in m80.parameters()) == {'11n': 2624080, '11l': 25372160}[variant]
        torch.save({'epoch': -1, 'model': m.half(), 'ema': None, 'train_args': {}}, p)
    finally:
        for n in _MODS:
            sys.modules.pop(n, None)
    assert 'ultralytics' not in sys.modules
    w = W.load_weights(p)
    assert w['variant'] == variant and w['nc'] == 5 and w['names'] == W.CLASS_NAMES
    assert set(w['state_dict']) == set(ours['state_dict'])
    for k, v in ours['state_dict'].items():
        assert torch.equal(w['state_dict'][k], v.half().float()), k


def test_yolo11_conv_gflops_match_published():
    """2 x conv MACs of the layer table at 640 x 640, nc = 80 (each layer at the stride its position in yolo11.yaml
    gives it) against the GFLOPs ultralytics publishes for the fused models: pins the strides / resolutions of the
    graph on top of the parameter counts."""
    def res_of(p):
        i = int(p.split('.')[1])
        if i == 23:
            return (80, 40, 20)[int(p.split('.')[3])]
        return {0: 320, 1: 160, 2: 160, 3: 80, 4: 80, 16: 80, 5: 40, 6: 40, 13: 40, 17: 40, 19: 40}.get(i, 20)
    for v, pub in (('11n', 6.5), ('11s', 21.5), ('11m', 68.0), ('11l', 86.9), ('11x', 194.9)):
        macs = sum(res_of(p) ** 2 * cout * (cin // g) * k * k for (p, cin, cout, k, g, bn) in W.conv_layers11(v, 80))
        assert abs(2 * macs / 1e9 - pub) < 0.06, (v, 2 * macs / 1e9, pub)


def test_yolov8_parameter_counts_and_gflops_match_published():
    """The yolov8 layer table (yolov8.yaml restated in weights.conv_bn_layers / the device model's graph builder / the
    oracle) against the numbers ultralytics publishes for nc = 80: parameter counts of the unfused models and GFLOPs
    of the fused models at 640 x 640 (2 x conv MACs at each layer's stride)."""
    want = {'n': (3157200, 8.7), 's': (11166560, 28.6), 'm': (25902640, 78.9), 'l': (43691520, 165.2),
            'x': (68229648, 257.8)}

    def res_of(p):
        i = int(p.split('.')[1])
        if i == 22:
            return (80, 40, 20)[int(p.split('.')[3])]
        return {0: 320, 1: 160, 2: 160, 3: 80, 4: 80, 15: 80, 5: 40, 6: 40, 12: 40, 16: 40, 18: 40}.get(i, 20)
    for v, (npar, gf) in want.items():
        L, cb, cc = W.conv_bn_layers(v, 80)
        params = 16 + sum(cout * cin * k * k + 2 * cout for (_, cin, cout, k) in L) + 3 * (64 * cb + 64 + 80 * cc + 80)
        macs = sum(res_of(p) ** 2 * cout * cin * k * k for (p, cin, cout, k) in L)
        macs += sum(r * r * (64 * cb + 80 * cc) for r in (80, 40, 20))
        assert params == npar, (v, params)
        assert abs(2 * macs / 1e9 - gf) < 0.06, (v, 2 * macs / 1e9)
