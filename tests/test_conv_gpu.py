"""tcgen05 implicit-GEMM conv vs a torch fp32 reference of the same op (bf16-rounded operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_case(B, H, W, cin, cout, k, s, act, res, out_f32, in_extra=0, out_extra=0, seed=0):
    from caesar_yolo_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device("cuda:0")
    in_ctot = cin + in_extra
    in_coff = in_extra
    x_full = (torch.randn(B, H, W, in_ctot, generator=g) * 1.0).to(torch.bfloat16).to(dev)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, generator=g) * 0.1
    Ho, Wo = H // s, W // s
    out_ctot = (cout + 7) // 8 * 8 + out_extra
    out_coff = out_extra
    out = torch.full((B, Ho, Wo, out_ctot), 7.0, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
    r = None
    if res:
        r = torch.randn(B, Ho, Wo, cout, generator=g).to(torch.bfloat16).to(dev)
    wp, bp = ops.pack_conv_weight(w, b, dev)
    ops.conv2d_nhwc(x_full, in_coff, cin, wp, bp, cout, k, s, out, out_coff, act=act, res=r, res_coff=0)
    torch.cuda.synchronize()
    # reference
    xr = x_full[..., in_coff:in_coff + cin].float().permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(xr, w.float().to(dev), b.to(dev), stride=s, padding=k // 2)
    if act:
        y = torch.nn.functional.silu(y)
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    y = y.permute(0, 2, 3, 1)
    got = out[..., out_coff:out_coff + cout].float()
    err = (got - y).abs().max().item()
    tol = 2e-3 if out_f32 else 3e-2
    assert err < tol * max(1.0, y.abs().max().item()), (err, y.abs().max().item())
    if out_extra:
        assert (out[..., :out_coff] == 7.0).all()  # untouched neighbouring slice


@pytest.mark.parametrize("case", [
    dict(B=2, H=16, W=16, cin=64, cout=64, k=1, s=1, act=True, res=False, out_f32=False),
    dict(B=2, H=16, W=16, cin=64, cout=64, k=3, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=20, W=20, cin=128, cout=128, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=2, H=40, W=40, cin=64, cout=128, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=1, H=80, W=40, cin=256, cout=256, k=3, s=1, act=True, res=False, out_f32=False, in_extra=64, out_extra=128),
    dict(B=5, H=20, W=10, cin=320, cout=512, k=1, s=1, act=True, res=False, out_f32=False),
    dict(B=2, H=20, W=20, cin=64, cout=64, k=1, s=1, act=False, res=False, out_f32=True, out_extra=16),
    dict(B=2, H=20, W=20, cin=256, cout=5, k=1, s=1, act=False, res=False, out_f32=True, out_extra=64),
    dict(B=2, H=32, W=32, cin=16, cout=16, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=2, H=32, W=32, cin=32, cout=32, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=1, H=160, W=160, cin=64, cout=64, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=4, H=20, W=20, cin=512, cout=512, k=3, s=1, act=True, res=False, out_f32=False),
    # 64-byte channel rows (kc = 32, SWIZZLE_64B) with the 3x3 tap reuse through descriptor offsets (yolo11 / v8n,s)
    dict(B=3, H=80, W=80, cin=32, cout=32, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=2, H=32, W=48, cin=96, cout=64, k=3, s=1, act=True, res=False, out_f32=False, in_extra=32, out_extra=8),
])
def test_conv_matches_torch(case):
    _run_case(**case)


# Kernel variants: mode 0 = one A box per tap, 1 = vertical tap reuse, 2 = full 3x3 halo reuse; halves = 128-row
# accumulators per work unit.  The planner picks these from the shape; the environment overrides exist for this test.
VARIANT_SHAPES = [
    dict(B=2, H=16, W=16, cin=64, cout=64, k=3, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=40, W=40, cin=128, cout=128, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=1, H=80, W=40, cin=256, cout=256, k=3, s=1, act=True, res=False, out_f32=False, in_extra=64, out_extra=128),
    dict(B=5, H=32, W=8, cin=64, cout=5, k=3, s=1, act=False, res=False, out_f32=True, out_extra=64),
    dict(B=7, H=20, W=20, cin=192, cout=256, k=1, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=40, W=40, cin=64, cout=128, k=3, s=2, act=True, res=False, out_f32=False),
]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("halves", [1, 2])
@pytest.mark.parametrize("shape", range(len(VARIANT_SHAPES)))
def test_conv_variants(monkeypatch, mode, halves, shape):
    monkeypatch.setenv("CY_CONV_MODE", str(mode))
    monkeypatch.setenv("CY_CONV_HALVES", str(halves))
    _run_case(seed=shape + 1, **VARIANT_SHAPES[shape])


def test_conv_large_batch_persistent():
    """More work units than SMs: every CTA loops over several units (ring phases, TMEM double buffering)."""
    _run_case(B=24, H=80, W=80, cin=128, cout=128, k=3, s=1, act=True, res=True, out_f32=False, seed=11)
    _run_case(B=32, H=40, W=40, cin=256, cout=512, k=1, s=1, act=True, res=False, out_f32=False, seed=12)
    _run_case(B=16, H=160, W=160, cin=64, cout=64, k=3, s=1, act=True, res=False, out_f32=False, seed=13)


PAIR_SHAPES = [
    dict(B=2, H=16, W=16, cin=64, cout=64, k=3, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=40, W=40, cin=128, cout=128, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=1, H=80, W=40, cin=256, cout=256, k=3, s=1, act=True, res=False, out_f32=False, in_extra=64, out_extra=128),
    dict(B=7, H=20, W=20, cin=192, cout=256, k=1, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=40, W=40, cin=64, cout=128, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=2, H=20, W=20, cin=64, cout=64, k=1, s=1, act=False, res=False, out_f32=True, out_extra=16),
]


@pytest.mark.parametrize("halves", [1, 2])
@pytest.mark.parametrize("shape", range(len(PAIR_SHAPES)))
def test_conv_cta_pairs(monkeypatch, halves, shape):
    """tcgen05 cta_group::2 path (cluster of two CTAs per work unit), forced on small shapes."""
    monkeypatch.setenv("CY_CONV_PAIR", "2")
    monkeypatch.setenv("CY_CONV_HALVES", str(halves))
    _run_case(seed=shape + 21, **PAIR_SHAPES[shape])


def test_conv_cta_pairs_persistent(monkeypatch):
    monkeypatch.setenv("CY_CONV_PAIR", "2")
    _run_case(B=24, H=80, W=80, cin=128, cout=128, k=3, s=1, act=True, res=True, out_f32=False, seed=31)
    _run_case(B=32, H=40, W=40, cin=256, cout=512, k=1, s=1, act=True, res=False, out_f32=False, seed=32)
    _run_case(B=16, H=160, W=160, cin=64, cout=64, k=3, s=1, act=True, res=False, out_f32=False, seed=33)


S2_SHAPES = [
    dict(B=2, H=64, W=32, cin=64, cout=128, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=3, H=80, W=80, cin=128, cout=256, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=1, H=96, W=48, cin=192, cout=64, k=3, s=2, act=True, res=False, out_f32=False, in_extra=64, out_extra=64),
]


@pytest.mark.parametrize("pair", [0, 2])
@pytest.mark.parametrize("halves", [1, 2])
@pytest.mark.parametrize("shape", range(len(S2_SHAPES)))
def test_conv_stride2_parity_patch_reuse(monkeypatch, pair, halves, shape):
    """MODE 3: stride-2 3x3 convs read four parity-class boxes per channel chunk instead of nine tap boxes."""
    monkeypatch.setenv("CY_CONV_PAIR", str(pair))
    monkeypatch.setenv("CY_CONV_HALVES", str(halves))
    _run_case(seed=shape + 41, **S2_SHAPES[shape])


WIDE_SHAPES = [
    dict(B=1, H=80, W=40, cin=256, cout=256, k=3, s=1, act=True, res=False, out_f32=False, in_extra=64, out_extra=128),
    dict(B=7, H=20, W=20, cin=192, cout=256, k=1, s=1, act=True, res=False, out_f32=False),
    dict(B=3, H=40, W=40, cin=128, cout=512, k=3, s=1, act=True, res=True, out_f32=False),
    dict(B=3, H=80, W=80, cin=128, cout=256, k=3, s=2, act=True, res=False, out_f32=False),
    dict(B=5, H=20, W=10, cin=320, cout=512, k=1, s=1, act=False, res=False, out_f32=False, out_extra=8),
]


@pytest.mark.parametrize("shape", range(len(WIDE_SHAPES)))
def test_conv_wide_pair_tiles(monkeypatch, shape):
    """256 x 256 pair units (cta_group::2, N = 256, one 128-row half per CTA, two 256-column accumulators in TMEM),
    forced on small shapes; the planner picks them for Cout % 256 == 0 pair layers."""
    from caesar_yolo_b200 import ops
    import ctypes
    monkeypatch.setenv("CY_CONV_PAIR", "2")
    monkeypatch.setenv("CY_CONV_WIDE", "2")
    c = WIDE_SHAPES[shape]
    info = (ctypes.c_int * 8)()
    ops.check(ops.lib.cy_conv_plan_info(c['B'], c['H'], c['W'], c['cin'], c['cout'], c['k'], c['s'], info))
    assert info[0] >= 10 and info[1] == 1 and info[3] == 256 and info[6] == 2   # pair, one half, N 256, 2 acc buffers
    _run_case(seed=shape + 51, **c)


def test_conv_wide_pair_tiles_persistent():
    """Planner-selected wide tiles with several units per CTA pair (ring phases, both TMEM buffers in use)."""
    from caesar_yolo_b200 import ops
    import ctypes
    info = (ctypes.c_int * 8)()
    ops.check(ops.lib.cy_conv_plan_info(32, 40, 40, 1024, 512, 1, 1, info))
    assert info[3] == 256 and info[0] >= 10
    _run_case(B=32, H=40, W=40, cin=1024, cout=512, k=1, s=1, act=True, res=False, out_f32=False, seed=61)
    ops.check(ops.lib.cy_conv_plan_info(64, 20, 20, 256, 256, 3, 1, info))
    assert info[3] == 256 and info[0] == 10     # 20x20 maps: one box per tap (mode 0), pair, wide
    _run_case(B=64, H=20, W=20, cin=256, cout=256, k=3, s=1, act=True, res=True, out_f32=False, seed=62)


@pytest.mark.parametrize("B,H,W,cout,act", [
    (2, 64, 64, 16, 1),       # v8n stem, one partly filled 64-column block
    (1, 640, 320, 64, 1),     # v8l stem on a 512 x 256 edge tile (Wo = 160: 2.5 column blocks)
    (3, 96, 160, 32, 1),
    (2, 32, 96, 48, 2),       # ex2 + rcp SiLU
    (2, 64, 32, 80, 0),       # v8x width, no activation
    (5, 320, 320, 64, 1),
])
def test_stem_conv_matches_torch(B, H, W, cout, act):
    """Fused mma.sync stem (model.0: 3x3 stride-2 pad-1 conv on the NHWC-4 model input + bias + SiLU) vs torch fp32 on
    the same bf16-rounded operands.  The 4th input channel must be ignored whatever it holds."""
    from caesar_yolo_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + cout)
    dev = torch.device("cuda:0")
    x = torch.randn(B, H, W, 4, generator=g).to(torch.bfloat16)
    x[..., 3] = 3.0   # garbage in the padding channel: its weights are zero
    w = (torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g) * 0.1
    got = ops.stem_conv(x.to(dev).contiguous(), w, b, act=act).float().cpu()
    y = torch.nn.functional.conv2d(x[..., :3].float().permute(0, 3, 1, 2), w, b, stride=2, padding=1)
    if act:
        y = torch.nn.functional.silu(y)
    y = y.permute(0, 2, 3, 1)
    assert got.shape == y.shape
    err = (got - y).abs().max().item()
    assert err < 2e-2 * max(1.0, y.abs().max().item()), (err, y.abs().max().item())
