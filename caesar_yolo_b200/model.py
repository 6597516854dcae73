"""`YOLO(weights)` — the object scripts/run.py:347 creates and the reference calls as
`model(image, save=False, device=..., imgsz=..., conf=..., iou=..., ...)` (caesar_yolo/evaluation.py:181-193).
Here it owns a DeviceModel (tcgen05 conv stack) and answers the same call with the same result shape:
an iterable of objects exposing `.boxes.xyxy/.conf/.cls` tensors (`.cpu().numpy()` works on them)."""
import numpy as np
import torch

from . import ops, weights as W


class _Boxes(object):
    def __init__(self, det):
        self.data = det
        self.xyxy = det[:, :4]
        self.conf = det[:, 4]
        self.cls = det[:, 5]

    def __len__(self):
        return int(self.data.shape[0])


class Results(object):
    def __init__(self, det, names, orig_shape):
        self.boxes = _Boxes(det)
        self.names = names
        self.orig_shape = orig_shape


class YOLO(object):
    def __init__(self, weights, device=None, precision=None):
        """weights: path of a caesar_yolo_b200 weight file (weights.py) or an already loaded weight dict.
        precision: 'fp16' / 'bf16' storage of weights and activations (None: ops.DEFAULT_PRECISION)."""
        self.precision = precision
        self.weights = W.load_weights(weights) if isinstance(weights, str) else weights
        self.names = dict(self.weights['names'])
        self._models = {}
        self._device = device

    def device_model(self):
        dev = torch.cuda.current_device()
        m = self._models.get(dev)
        if m is None:
            m = ops.DeviceModel(self.weights, precision=self.precision)
            self._models[dev] = m
        return m

    def __call__(self, image, save=False, device=None, imgsz=640, conf=0.25, iou=0.7, **kwargs):
        """image: H x W x 3 array (already preprocessed, any float dtype) -> [Results]."""
        if device is not None and str(device) not in ('cpu', ''):
            d = str(device)
            torch.cuda.set_device(int(d.split(':')[1]) if ':' in d else int(d))
        elif str(device) == 'cpu':
            raise ops.CaesarB200Error("device 'cpu' requested: the B200 build has no CPU path (use --devices=cuda:N)")
        a = np.asarray(image)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("expected an H x W x 3 image")
        dev = torch.device('cuda:%d' % torch.cuda.current_device())
        H, Wd = a.shape[:2]
        chain = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev).unsqueeze(0)
        m = self.device_model()
        x, _ = ops.letterbox_resize(chain, imgsz, dtype=m.dtype)
        Sh, Sw, lb = ops.letterbox_shape(H, Wd, imgsz)
        heads = m.forward(x)
        dets, nd = ops.postprocess(heads, 1, Sh, Sw, m.nc, conf, iou, ops.letterbox_array([lb], dev), dev)
        n = int(nd[0].item())
        return [Results(dets[0, :n].clone(), self.names, (H, Wd))]
