"""Raw (non-autograd) launchers: torch tensors in, C-ABI call out.  Plumbing only."""
import ctypes

import torch

from . import lib as _lib
from .lib import (CB_ACT_NONE, CB_BF16, CB_F16, CB_F32, CB_MAJOR_K, CB_MAJOR_MN, GemmDesc)

_DT = {torch.float16: CB_F16, torch.bfloat16: CB_BF16, torch.float32: CB_F32}


def dt(t):
    return _DT[t.dtype]


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(A, B, D, *, M, N, K, batch=1, lda=None, ldb=None, ldd=None, a_bs=0, b_bs=0, d_bs=0,
         a_major=CB_MAJOR_K, b_major=CB_MAJOR_K, bias=None, bias_row_div=0, ldbias=0,
         R=None, ldr=0, r_bs=0, alpha=1.0, act=CB_ACT_NONE, d_transposed=False, conv=None):
    """Launch cb_gemm.  A/B/D/R are CUDA tensors used as raw storage; all strides in elements.

    conv: None or dict(img_n,img_h,img_w,out_h,out_w,kh,kw,stride,pad_top,pad_left,b_tap_rows,flip_taps).
    """
    L = _lib.load()
    d = GemmDesc()
    d.M, d.N, d.K, d.batch = int(M), int(N), int(K), int(batch)
    d.ab_dtype = dt(A)
    assert A.dtype == B.dtype
    d.A, d.lda, d.a_batch_stride, d.a_major = A.data_ptr(), int(lda), int(a_bs), a_major
    d.B, d.ldb, d.b_batch_stride, d.b_major = B.data_ptr(), int(ldb), int(b_bs), b_major
    if conv is not None:
        d.conv = 1
        for k, v in conv.items():
            setattr(d, k, int(v))
    d.D, d.d_dtype, d.ldd, d.d_batch_stride = D.data_ptr(), dt(D), int(ldd), int(d_bs)
    d.d_transposed = 1 if d_transposed else 0
    if bias is not None:
        assert bias.dtype == torch.float32
        d.bias, d.bias_row_div, d.ldbias = bias.data_ptr(), int(bias_row_div), int(ldbias)
    if R is not None:
        d.R, d.r_dtype, d.ldr, d.r_batch_stride = R.data_ptr(), dt(R), int(ldr), int(r_bs)
    d.alpha, d.act = float(alpha), int(act)
    from . import ops as _ops
    ws = _ops._splitk_workspace(torch.cuda.current_device())
    d.splitk_ws, d.splitk_ws_bytes = ws.data_ptr(), ws.numel()
    _lib.check(L.cb_gemm(ctypes.byref(d), stream_ptr()), "cb_gemm")
