"""ctypes binding of include/mcb200.h.  No fallback: a missing library or a failing call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcb200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libmcb200.so not built: run `python __graft_entry__.py` (nvcc, sm_100a). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)
lib.mcb_last_error.restype = C.c_char_p
lib.mcb_version.restype = C.c_int

vp, ci, fp = C.c_void_p, C.c_int, C.POINTER(C.c_float)


class ConvFwdArgs(C.Structure):
    _fields_ = [("x", vp * 2), ("cin", ci * 2), ("n", ci), ("h", ci), ("w", ci), ("weight", vp), ("cout", ci),
                ("ksize", ci), ("stride", ci), ("bias", vp), ("relu", ci), ("stats", vp), ("y", vp)]


class ConvDgradArgs(C.Structure):
    _fields_ = [("dy", vp), ("n", ci), ("h", ci), ("w", ci), ("weight", vp), ("cout", ci), ("cin_total", ci),
                ("ci_off", ci), ("cin", ci), ("ksize", ci), ("stride", ci), ("dx", vp), ("relu_mask", vp),
                ("accumulate", ci)]


class ConvWgradArgs(C.Structure):
    _fields_ = [("dy", vp), ("x", vp), ("n", ci), ("h", ci), ("w", ci), ("cout", ci), ("cin_total", ci),
                ("ci_off", ci), ("cin", ci), ("ksize", ci), ("stride", ci), ("dw", vp)]


class ConvtFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("n", ci), ("h", ci), ("w", ci), ("cin", ci), ("weight", vp), ("cout", ci), ("bias", vp),
                ("relu", ci), ("y", vp)]


class ConvtDgradArgs(C.Structure):
    _fields_ = [("dy", vp), ("n", ci), ("h", ci), ("w", ci), ("cin", ci), ("weight", vp), ("cout", ci), ("dx", vp),
                ("relu_mask", vp), ("accumulate", ci)]


class ConvtWgradArgs(C.Structure):
    _fields_ = [("dy", vp), ("x", vp), ("n", ci), ("h", ci), ("w", ci), ("cin", ci), ("cout", ci), ("dw", vp)]


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libmcb200 %s failed (%d): %s" % (what, rc, lib.mcb_last_error().decode()))


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, args=None, *extra):
    fn = getattr(lib, name)
    if args is None:
        rc = fn(*extra, stream_ptr())
    else:
        rc = fn(C.byref(args), *extra, stream_ptr())
    check(rc, name)
