"""Restatement of the ultralytics YOLO11 detection model (cfg/models/11/yolo11.yaml: C3k2 / C3k / Bottleneck, SPPF,
C2PSA with PSABlock + Attention, Detect with depthwise-separable class branch), torch CPU fp32.  ultralytics is not
installed here, so this is written from the published architecture; the layer table it shares with
caesar_yolo_b200/weights.py reproduces the published parameter counts of all five scales exactly
(tests/test_weights_cpu.py).  The reference reaches it through `YOLO(weights)` / `model(image, ...)`
(scripts/run.py:347, caesar_yolo/evaluation.py:181-193; README.md:200-207 lists yolo11 weights).  Oracle; test-only.
"""
import math

import torch
import torch.nn.functional as F

from .yolo import OracleYolo

SCALES11 = {'n': (0.50, 0.25, 1024), 's': (0.50, 0.50, 1024), 'm': (0.50, 1.00, 512), 'l': (1.00, 1.00, 512),
            'x': (1.00, 1.50, 512)}


def arch11(variant):
    depth, width, maxc = SCALES11[variant[2:]]
    ch = lambda c: int(math.ceil(min(c, maxc) * width / 8) * 8)
    return dict(c64=ch(64), c128=ch(128), c256=ch(256), c512=ch(512), c1024=ch(1024), n=max(round(2 * depth), 1),
                c3k=variant[2:] in 'mlx')


class OracleYolo11(OracleYolo):
    """Same conventions as OracleYolo: Conv + BN fused; emulate_bf16 rounds weights and every stored layer output."""

    def __init__(self, weights, emulate_bf16=False):
        OracleYolo.__init__(self, weights, emulate_bf16)
        self.a = arch11(self.variant)

    def conv(self, x, p, k, s, act=True):
        w, b = self.w[p]
        y = F.conv2d(x, w, b, stride=s, padding=k // 2, groups=x.shape[1] // w.shape[1])
        if act:
            y = F.silu(y)
        return y

    def bottleneck(self, x, p, k=3):
        """Bottleneck with shortcut: x + cv2(cv1(x)) (cv1/cv2 3x3 Conv-BN-SiLU)."""
        t = self.cbs(x, p + '.cv1', k)
        return self._q(x + self.conv(t, p + '.cv2', k, 1))

    def c3k(self, x, p):
        """C3k(c, c, n=2): cv3(cat(m(cv1(x)), cv2(x))), m = two Bottleneck(c/2, c/2, e=1.0)."""
        a = self.cbs(x, p + '.cv1')
        b = self.cbs(x, p + '.cv2')
        for j in range(2):
            a = self.bottleneck(a, '%s.m.%d' % (p, j))
        return self.cbs(torch.cat([a, b], 1), p + '.cv3')

    def c3k2(self, x, p, c3k):
        y = self.cbs(x, p + '.cv1')
        c = y.shape[1] // 2
        ys = [y[:, :c], y[:, c:]]
        for i in range(self.a['n']):
            m = '%s.m.%d' % (p, i)
            ys.append(self.c3k(ys[-1], m) if c3k else self.bottleneck(ys[-1], m))
        return self.cbs(torch.cat(ys, 1), p + '.cv2')

    def attention(self, x, p):
        """Attention(dim, num_heads = dim / 64, attn_ratio = 0.5): per head q, k (32) and v (64) are consecutive
        channel groups of the qkv conv output; softmax(q^T k / sqrt(32)) over the keys; + depthwise 3x3 `pe` on v;
        1x1 `proj`.  Returns proj(...) WITHOUT the residual (the caller adds it before rounding)."""
        B, C, H, W = x.shape
        N = H * W
        nh = C // 64
        kd, hd = 32, 64
        qkv = self._q(self.conv(x, p + '.qkv', 1, 1, act=False)).view(B, nh, 2 * kd + hd, N)
        q, k, v = qkv[:, :, :kd], qkv[:, :, kd:2 * kd], qkv[:, :, 2 * kd:]
        attn = (q.transpose(-2, -1) @ k) * (kd ** -0.5)
        attn = attn.softmax(dim=-1)
        o = self._q((v @ attn.transpose(-2, -1)).reshape(B, C, H, W))
        o = self._q(o + self.conv(v.reshape(B, C, H, W), p + '.pe', 3, 1, act=False))
        return self.conv(o, p + '.proj', 1, 1, act=False)

    def c2psa(self, x, p):
        y = self.cbs(x, p + '.cv1')
        c = y.shape[1] // 2
        a, b = y[:, :c], y[:, c:]
        for i in range(self.a['n']):
            m = '%s.m.%d' % (p, i)
            b = self._q(b + self.attention(b, m + '.attn'))
            t = self.cbs(b, m + '.ffn.0')
            b = self._q(b + self.conv(t, m + '.ffn.1', 1, 1, act=False))
        return self.cbs(torch.cat([a, b], 1), p + '.cv2')

    def forward_heads(self, x):
        big = self.a['c3k']
        up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest')
        x = self._q(x)
        x0 = self.cbs(x, 'model.0', 3, 2)
        x1 = self.cbs(x0, 'model.1', 3, 2)
        x2 = self.c3k2(x1, 'model.2', big)
        x3 = self.cbs(x2, 'model.3', 3, 2)
        x4 = self.c3k2(x3, 'model.4', big)
        x5 = self.cbs(x4, 'model.5', 3, 2)
        x6 = self.c3k2(x5, 'model.6', True)
        x7 = self.cbs(x6, 'model.7', 3, 2)
        x8 = self.c3k2(x7, 'model.8', True)
        y = self.cbs(x8, 'model.9.cv1')
        p1 = F.max_pool2d(y, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        x9 = self.cbs(torch.cat([y, p1, p2, p3], 1), 'model.9.cv2')
        x10 = self.c2psa(x9, 'model.10')
        x13 = self.c3k2(torch.cat([up(x10), x6], 1), 'model.13', big)
        x16 = self.c3k2(torch.cat([up(x13), x4], 1), 'model.16', big)
        x17 = self.cbs(x16, 'model.17', 3, 2)
        x19 = self.c3k2(torch.cat([x17, x13], 1), 'model.19', big)
        x20 = self.cbs(x19, 'model.20', 3, 2)
        x22 = self.c3k2(torch.cat([x20, x10], 1), 'model.22', True)
        outs = []
        for l, f in enumerate((x16, x19, x22)):
            b = self.cbs(f, 'model.23.cv2.%d.0' % l, 3)
            b = self.cbs(b, 'model.23.cv2.%d.1' % l, 3)
            b = self.conv(b, 'model.23.cv2.%d.2' % l, 1, 1, act=False)
            c = self.cbs(f, 'model.23.cv3.%d.0.0' % l, 3)
            c = self.cbs(c, 'model.23.cv3.%d.0.1' % l, 1)
            c = self.cbs(c, 'model.23.cv3.%d.1.0' % l, 3)
            c = self.cbs(c, 'model.23.cv3.%d.1.1' % l, 1)
            c = self.conv(c, 'model.23.cv3.%d.2' % l, 1, 1, act=False)
            outs.append(torch.cat([b, c], 1))
        return outs
