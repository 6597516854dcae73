"""ctypes binding of libcaesar_b200.so (include/caesar_b200.h).  No fallback: if the library is missing the
import of any product module fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcaesar_b200.so")


class CaesarB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise CaesarB200Error(
            "libcaesar_b200.so not built (%s). Run `python -m caesar_yolo_b200.build` — there is no CPU fallback."
            % LIB_PATH)
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()
lib.cy_last_error.restype = ctypes.c_char_p

c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_uptr = ctypes.c_size_t


def check(rc):
    if rc != 0:
        raise CaesarB200Error("libcaesar_b200 error %d: %s" % (rc, lib.cy_last_error().decode()))
    return rc


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_uptr(torch.cuda.current_stream().cuda_stream)
