// Host-side letterbox geometry (ultralytics LetterBox(auto=True, scaleup=True, center=True, stride=32) and
// ops.scale_boxes; reference call site caesar_yolo/evaluation.py:181-193).
#include "common.h"
#include <math.h>

// Python round(): half to even == nearbyint in the default rounding mode.
static inline int py_round(double x) { return (int)nearbyint(x); }

extern "C" int cy_letterbox_shape(int Ty, int Tx, int imgsz, int* Sh_host, int* Sw_host, cy_letterbox* lb_host) {
    if (Ty <= 0 || Tx <= 0 || imgsz <= 0 || imgsz % 32) return cy::set_error(CY_ERR_INVALID, "invalid letterbox arguments");
    const double r = fmin((double)imgsz / Ty, (double)imgsz / Tx);
    const int new_w = py_round(Tx * r), new_h = py_round(Ty * r);
    double dw = (double)((imgsz - new_w) % 32), dh = (double)((imgsz - new_h) % 32);
    dw /= 2;
    dh /= 2;
    const int top = py_round(dh - 0.1), bottom = py_round(dh + 0.1);
    const int left = py_round(dw - 0.1), right = py_round(dw + 0.1);
    const int Sh = new_h + top + bottom, Sw = new_w + left + right;
    if (Sh_host) *Sh_host = Sh;
    if (Sw_host) *Sw_host = Sw;
    if (lb_host) {
        const double gain = fmin((double)Sh / Ty, (double)Sw / Tx);
        lb_host->gain = (float)gain;
        lb_host->pad_x = (float)py_round((Sw - Tx * gain) / 2 - 0.1);
        lb_host->pad_y = (float)py_round((Sh - Ty * gain) / 2 - 0.1);
        lb_host->w0 = Tx;
        lb_host->h0 = Ty;
    }
    return CY_OK;
}
