"""Restatement of the ultralytics pieces the reference calls through `model(image, ...)`
(caesar_yolo/evaluation.py:181-193): LetterBox + predictor preprocess, YOLOv8 DetectionModel forward,
Detect decode, non_max_suppression (torchvision.ops.nms), scale_boxes/clip_boxes.  ultralytics is not
installed here; see SURVEY.md Appendix A.4-A.6.  torch CPU fp32.  Oracle; test-only.
"""
import math

import cv2
import numpy as np
import torch
import torch.nn.functional as F
import torchvision

CLASS_NAMES = {0: 'spurious', 1: 'compact', 2: 'extended', 3: 'extended-multisland', 4: 'flagged'}
SCALES = {'n': (0.33, 0.25, 1024), 's': (0.33, 0.50, 1024), 'm': (0.67, 0.75, 768), 'l': (1.00, 1.00, 512),
          'x': (1.00, 1.25, 512)}


def make_divisible(x, d):
    return int(math.ceil(x / d) * d)


def arch(variant):
    depth, width, maxc = SCALES[variant]
    ch = lambda c: make_divisible(min(c, maxc) * width, 8)
    rep = lambda n: max(round(n * depth), 1)
    return dict(c1=ch(64), c2=ch(128), c3=ch(256), c4=ch(512), c5=ch(1024), n2=rep(3), n4=rep(6), n6=rep(6),
                n8=rep(3), nh=rep(3))


# ------------------------------------------------------------------------------------------------ forward

def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fp16(x):
    return x.to(torch.float16).to(torch.float32)


def _tf32(x):
    """Round to TF32 (10 explicit mantissa bits, round-to-nearest, ties away: cvt.rna.tf32.f32) — what cuDNN's default
    `allow_tf32` convolution path of the reference's own --devices=cuda run feeds the tensor cores."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class OracleYolo(object):
    """YOLOv8 DetectionModel with Conv+BN fused (ultralytics fuse()), fp32 on CPU.  emulate_bf16=True rounds
    weights and every layer output to bf16 (what the B200 path stores) to separate quantisation from kernel
    bugs; the default is the reference's fp32."""

    def __init__(self, weights, emulate_bf16=False):
        # emulate_bf16: False (fp32), True / 'bf16' (weights + every stored activation rounded to bf16), 'tf32' (conv
        # operands rounded to TF32, fp32 storage: the reference's own GPU arithmetic), 'perm' (fp32 with the input
        # channels of every conv visited in reverse order: same math, different accumulation order)
        self.variant = weights['variant']
        self.nc = weights['nc']
        self.names = dict(weights['names'])
        self.a = arch(self.variant) if not str(self.variant).startswith('11') else None
        self.mode = emulate_bf16 if isinstance(emulate_bf16, str) else ('bf16' if emulate_bf16 else 'fp32')
        self.emu = self.mode in ('bf16', 'fp16')
        self._rnd = _fp16 if self.mode in ('fp16', 'w16') else _bf16
        wq = self.emu or self.mode == 'w16'      # 'w16': fp16-rounded weights, fp32 activations (diagnostic)
        sd = weights['state_dict']
        self.w = {}
        for k in sd:
            if k.endswith('.conv.weight') and '.dfl.' not in k:
                p = k[:-len('.conv.weight')]
                w = sd[k].float()
                gamma, beta = sd[p + '.bn.weight'].float(), sd[p + '.bn.bias'].float()
                mean, var = sd[p + '.bn.running_mean'].float(), sd[p + '.bn.running_var'].float()
                s = gamma / torch.sqrt(var + 1e-3)
                wf = w * s.view(-1, 1, 1, 1)
                bf = beta - mean * s
                self.w[p] = (self._rnd(wf) if wq else wf, bf)
            elif k.endswith('.2.weight') and '.cv' in k:
                p = k[:-len('.weight')]
                self.w[p] = (self._rnd(sd[k].float()) if wq else sd[k].float(), sd[p + '.bias'].float())

    def _q(self, x):
        return self._rnd(x) if self.emu else x

    def conv(self, x, p, k, s, act=True):
        w, b = self.w[p]
        if self.mode == 'tf32':
            x, w = _tf32(x), _tf32(w)
        elif self.mode == 'perm':
            x, w = x.flip(1).contiguous(), w.flip(1).contiguous()
        y = F.conv2d(x, w, b, stride=s, padding=k // 2)
        if act:
            y = F.silu(y)
        return y

    def cbs(self, x, p, k=1, s=1):
        return self._q(self.conv(x, p, k, s))

    def c2f(self, x, p, n, shortcut):
        y = self.cbs(x, p + '.cv1')
        c = y.shape[1] // 2
        ys = [y[:, :c], y[:, c:]]
        for i in range(n):
            t = self.cbs(ys[-1], '%s.m.%d.cv1' % (p, i), 3)
            t = self.conv(t, '%s.m.%d.cv2' % (p, i), 3, 1)
            if shortcut:
                t = ys[-1] + t
            ys.append(self._q(t))
        return self.cbs(torch.cat(ys, 1), p + '.cv2')

    def forward_heads(self, x):
        """x: [B,3,H,W] fp32 -> list of 3 raw head maps [B, 64+nc, h, w]."""
        a = self.a
        x = self._q(x)
        x0 = self.cbs(x, 'model.0', 3, 2)
        x1 = self.cbs(x0, 'model.1', 3, 2)
        x2 = self.c2f(x1, 'model.2', a['n2'], True)
        x3 = self.cbs(x2, 'model.3', 3, 2)
        x4 = self.c2f(x3, 'model.4', a['n4'], True)
        x5 = self.cbs(x4, 'model.5', 3, 2)
        x6 = self.c2f(x5, 'model.6', a['n6'], True)
        x7 = self.cbs(x6, 'model.7', 3, 2)
        x8 = self.c2f(x7, 'model.8', a['n8'], True)
        y = self.cbs(x8, 'model.9.cv1')
        p1 = F.max_pool2d(y, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        x9 = self.cbs(torch.cat([y, p1, p2, p3], 1), 'model.9.cv2')
        x11 = torch.cat([F.interpolate(x9, scale_factor=2, mode='nearest'), x6], 1)
        x12 = self.c2f(x11, 'model.12', a['nh'], False)
        x14 = torch.cat([F.interpolate(x12, scale_factor=2, mode='nearest'), x4], 1)
        x15 = self.c2f(x14, 'model.15', a['nh'], False)
        x16 = self.cbs(x15, 'model.16', 3, 2)
        x18 = self.c2f(torch.cat([x16, x12], 1), 'model.18', a['nh'], False)
        x19 = self.cbs(x18, 'model.19', 3, 2)
        x21 = self.c2f(torch.cat([x19, x9], 1), 'model.21', a['nh'], False)
        outs = []
        for l, f in enumerate((x15, x18, x21)):
            b = self.cbs(f, 'model.22.cv2.%d.0' % l, 3)
            b = self.cbs(b, 'model.22.cv2.%d.1' % l, 3)
            b = self.conv(b, 'model.22.cv2.%d.2' % l, 1, 1, act=False)
            c = self.cbs(f, 'model.22.cv3.%d.0' % l, 3)
            c = self.cbs(c, 'model.22.cv3.%d.1' % l, 3)
            c = self.conv(c, 'model.22.cv3.%d.2' % l, 1, 1, act=False)
            outs.append(torch.cat([b, c], 1))
        return outs

    def decode(self, heads):
        """Detect._inference (App. A.6): DFL + dist2bbox(xywh) * stride, class sigmoid -> [B, 4+nc, A]."""
        B = heads[0].shape[0]
        no = 64 + self.nc
        xs, anchors, strides = [], [], []
        for h, s in zip(heads, (8., 16., 32.)):
            _, _, hh, ww = h.shape
            xs.append(h.reshape(B, no, hh * ww))
            sx = torch.arange(ww, dtype=torch.float32) + 0.5
            sy = torch.arange(hh, dtype=torch.float32) + 0.5
            gy, gx = torch.meshgrid(sy, sx, indexing='ij')
            anchors.append(torch.stack((gx, gy), -1).view(-1, 2))
            strides.append(torch.full((hh * ww, 1), s))
        x = torch.cat(xs, 2)
        anc = torch.cat(anchors).transpose(0, 1)  # [2,A]
        st = torch.cat(strides).transpose(0, 1)   # [1,A]
        box, cls = x[:, :64], x[:, 64:]
        A = box.shape[2]
        dist = (box.view(B, 4, 16, A).transpose(2, 1).softmax(1) *
                torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)).sum(1)  # [B,4,A]
        lt, rb = dist[:, :2], dist[:, 2:]
        x1y1 = anc.unsqueeze(0) - lt
        x2y2 = anc.unsqueeze(0) + rb
        cxy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        dbox = torch.cat((cxy, wh), 1) * st
        return torch.cat((dbox, cls.sigmoid()), 1)

    def forward(self, x):
        with torch.no_grad():
            return self.decode(self.forward_heads(x))


# ------------------------------------------------------------------------------------------------ pre / post

def letterbox(img, new_shape=(640, 640), stride=32):
    """ultralytics LetterBox(auto=True, scaleup=True, center=True) on an HWC numpy image (App. A.4)."""
    shape = img.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    return img


def preprocess(img_hwc, imgsz):
    """predictor.preprocess for one numpy image: letterbox, BGR->RGB flip, HWC->CHW, fp32, /255."""
    im = letterbox(img_hwc, (imgsz, imgsz))
    im = im[..., ::-1].transpose(2, 0, 1)
    im = np.ascontiguousarray(im)
    t = torch.from_numpy(im).unsqueeze(0).float()
    t /= 255
    return t


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms_single(pred, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680):
    """ultralytics ops.non_max_suppression for one image, agnostic=False, multi_label=False (App. A.6).
    pred: [4+nc, A] -> [N,6] (x1,y1,x2,y2,conf,cls) in descending score order."""
    nc = pred.shape[0] - 4
    xc = pred[4:4 + nc].amax(0) > conf_thres
    x = pred.transpose(0, 1)
    x = torch.cat((xywh2xyxy(x[:, :4]), x[:, 4:]), 1)
    x = x[xc]
    if not x.shape[0]:
        return torch.zeros((0, 6))
    box, cls = x[:, :4], x[:, 4:4 + nc]
    conf, j = cls.max(1, keepdim=True)
    x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
    n = x.shape[0]
    if not n:
        return torch.zeros((0, 6))
    if n > max_nms:
        x = x[x[:, 4].argsort(descending=True)[:max_nms]]
    c = x[:, 5:6] * max_wh
    scores = x[:, 4]
    boxes = x[:, :4] + c
    i = torchvision.ops.nms(boxes, scores, iou_thres)
    i = i[:max_det]
    return x[i]


def scale_boxes(img1_shape, boxes, img0_shape):
    """ultralytics ops.scale_boxes + clip_boxes (App. A.6)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
           round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    boxes = boxes.clone()
    boxes[..., 0] -= pad[0]
    boxes[..., 1] -= pad[1]
    boxes[..., 2] -= pad[0]
    boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, img0_shape[1])
    boxes[..., 1].clamp_(0, img0_shape[0])
    boxes[..., 2].clamp_(0, img0_shape[1])
    boxes[..., 3].clamp_(0, img0_shape[0])
    return boxes


class _Boxes(object):
    def __init__(self, det):
        self.xyxy = det[:, :4]
        self.conf = det[:, 4]
        self.cls = det[:, 5]


class _Result(object):
    def __init__(self, det):
        self.boxes = _Boxes(det)


class OracleModel(object):
    """Duck-typed stand-in for ultralytics.YOLO as the reference uses it (evaluation.py:46-47,181-193,261-264)."""

    def __init__(self, weights, emulate_bf16=False):
        if str(weights['variant']).startswith('11'):
            from .yolo11 import OracleYolo11
            self.net = OracleYolo11(weights, emulate_bf16)
        else:
            self.net = OracleYolo(weights, emulate_bf16)
        self.names = self.net.names

    def __call__(self, image, save=False, device='cpu', imgsz=640, conf=0.25, iou=0.7, **kw):
        im = preprocess(np.asarray(image), imgsz)
        pred = self.net.forward(im)[0]
        det = nms_single(pred, conf, iou)
        if det.shape[0]:
            det[:, :4] = scale_boxes(im.shape[2:], det[:, :4], image.shape[:2])
        return [_Result(det)]
