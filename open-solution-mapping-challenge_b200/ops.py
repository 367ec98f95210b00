"""Tensor-level wrappers over the C ABI (one function per entry point of include/mcb200.h).

All activations are torch CUDA tensors, NHWC (shape (N, H, W, C)), bf16, contiguous.  Conv weights are "packed":
bf16 (k*k, cout, cin).  These wrappers only marshal pointers and shapes; the arithmetic is in libmcb200.so."""
import ctypes as C

import torch

from . import _lib as L


def _chk(t, dtype=torch.bfloat16):
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.device, t.is_contiguous(), t.dtype)
    return t


def pack_conv_weight(w):
    """(cout, cin, kh, kw) -> (kh*kw, cout, cin), the library's tap-major layout"""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous()


def unpack_conv_weight(wp, k):
    t, co, ci = wp.shape
    return wp.reshape(k, k, co, ci).permute(2, 3, 0, 1).contiguous()


def pack_convt_weight(w):
    """ConvTranspose2d weight (cin, cout, kh, kw) -> (kh*kw, cout, cin)"""
    ci, co, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw, co, ci).contiguous()


def unpack_convt_weight(wp, k=4):
    t, co, ci = wp.shape
    return wp.reshape(k, k, co, ci).permute(3, 2, 0, 1).contiguous()


def conv_fwd(x, w, ksize, stride=1, bias=None, relu=False, stats=None, x2=None, out=None):
    _chk(x); _chk(w)
    n, h, wd, c0 = x.shape
    c1 = 0
    if x2 is not None:
        _chk(x2)
        assert x2.shape[:3] == x.shape[:3]
        c1 = x2.shape[3]
    cout = w.shape[1]
    assert w.shape == (ksize * ksize, cout, c0 + c1), (w.shape, ksize, cout, c0, c1)
    if out is None:
        out = torch.empty((n, h // stride, wd // stride, cout), dtype=torch.bfloat16, device=x.device)
    a = L.ConvFwdArgs()
    a.x[0] = x.data_ptr(); a.x[1] = x2.data_ptr() if x2 is not None else None
    a.cin[0] = c0; a.cin[1] = c1
    a.n, a.h, a.w = n, h, wd
    a.weight = w.data_ptr(); a.cout = cout; a.ksize = ksize; a.stride = stride
    a.bias = _chk(bias, torch.float32).data_ptr() if bias is not None else None
    a.relu = int(relu)
    a.stats = _chk(stats, torch.float32).data_ptr() if stats is not None else None
    a.y = _chk(out).data_ptr()
    L.call("mcb_conv_fwd", a)
    return out


def conv_dgrad(dy, w, ksize, stride, in_hw, cin=None, ci_off=0, relu_mask=None, accumulate=False, out=None):
    _chk(dy); _chk(w)
    n = dy.shape[0]
    h, wd = in_hw
    cout, cin_total = w.shape[1], w.shape[2]
    cin = cin_total if cin is None else cin
    if out is None:
        assert not accumulate
        out = torch.empty((n, h, wd, cin), dtype=torch.bfloat16, device=dy.device)
    a = L.ConvDgradArgs()
    a.dy = dy.data_ptr(); a.n, a.h, a.w = n, h, wd
    a.weight = w.data_ptr(); a.cout = cout; a.cin_total = cin_total; a.ci_off = ci_off; a.cin = cin
    a.ksize = ksize; a.stride = stride
    a.dx = _chk(out).data_ptr()
    a.relu_mask = _chk(relu_mask).data_ptr() if relu_mask is not None else None
    a.accumulate = int(accumulate)
    L.call("mcb_conv_dgrad", a)
    return out


def conv_wgrad(dy, x, dw, ksize, stride, ci_off=0):
    """dw (fp32, (k*k, cout, cin_total)) += wgrad"""
    _chk(dy); _chk(x); _chk(dw, torch.float32)
    n, h, wd, cin = x.shape
    a = L.ConvWgradArgs()
    a.dy = dy.data_ptr(); a.x = x.data_ptr(); a.n, a.h, a.w = n, h, wd
    a.cout = dw.shape[1]; a.cin_total = dw.shape[2]; a.ci_off = ci_off; a.cin = cin
    a.ksize = ksize; a.stride = stride; a.dw = dw.data_ptr()
    L.call("mcb_conv_wgrad", a)
    return dw


def convt_fwd(x, w, bias=None, relu=False, out=None):
    _chk(x); _chk(w)
    n, h, wd, cin = x.shape
    cout = w.shape[1]
    assert w.shape == (16, cout, cin)
    if out is None:
        out = torch.empty((n, 2 * h, 2 * wd, cout), dtype=torch.bfloat16, device=x.device)
    a = L.ConvtFwdArgs()
    a.x = x.data_ptr(); a.n, a.h, a.w, a.cin = n, h, wd, cin
    a.weight = w.data_ptr(); a.cout = cout
    a.bias = _chk(bias, torch.float32).data_ptr() if bias is not None else None
    a.relu = int(relu); a.y = _chk(out).data_ptr()
    L.call("mcb_convt_fwd", a)
    return out


def convt_dgrad(dy, w, relu_mask=None, accumulate=False, out=None):
    _chk(dy); _chk(w)
    n, h2, w2, cout = dy.shape
    h, wd = h2 // 2, w2 // 2
    cin = w.shape[2]
    if out is None:
        assert not accumulate
        out = torch.empty((n, h, wd, cin), dtype=torch.bfloat16, device=dy.device)
    a = L.ConvtDgradArgs()
    a.dy = dy.data_ptr(); a.n, a.h, a.w, a.cin = n, h, wd, cin
    a.weight = w.data_ptr(); a.cout = cout; a.dx = _chk(out).data_ptr()
    a.relu_mask = _chk(relu_mask).data_ptr() if relu_mask is not None else None
    a.accumulate = int(accumulate)
    L.call("mcb_convt_dgrad", a)
    return out


def convt_wgrad(dy, x, dw):
    _chk(dy); _chk(x); _chk(dw, torch.float32)
    n, h, wd, cin = x.shape
    a = L.ConvtWgradArgs()
    a.dy = dy.data_ptr(); a.x = x.data_ptr(); a.n, a.h, a.w, a.cin = n, h, wd, cin
    a.cout = dw.shape[1]; a.dw = dw.data_ptr()
    L.call("mcb_convt_wgrad", a)
    return dw
