"""Helpers for the -m gpu tests: call the C ABI with torch CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from b200vgan import lib

L = None


def L_():
    global L
    if L is None:
        L = lib.load()
    return L


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype).contiguous()


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def activation1d(x, la, lb, dtype=torch.float32):
    xt, y = dev(x, dtype), None
    y = torch.empty_like(xt)
    B, Cc, T = xt.shape
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    lat, lbt = dev(la), dev(lb)      # keep the parameter tensors alive across the call
    lib.check(L_().bvg_activation1d(ptr(xt), ptr(y), ptr(lat), ptr(lbt), B, Cc, T, code, stream()))
    torch.cuda.synchronize()
    return y.float().cpu().numpy()


def activation1d_packed(x, la, lb, mode):
    xt, lat, lbt = dev(x), dev(la), dev(lb)
    y = torch.empty_like(xt)
    B, Cc, T = xt.shape
    lib.check(L_().bvg_activation1d_packed(ptr(xt), ptr(y), ptr(lat), ptr(lbt), B, Cc, T, mode, stream()))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def conv1d(x, w, b, res, k, d, mode):
    xt, wt = dev(x), dev(w)
    bt = None if b is None else dev(b)
    rt = None if res is None else dev(res)
    B, Cin, T = xt.shape
    Cout = wt.shape[0]
    y = torch.empty(B, Cout, T, device="cuda")
    lib.check(L_().bvg_conv1d(ptr(xt), ptr(wt), ptr(bt), ptr(rt), ptr(y), B, Cin, Cout, T, k, d, mode, stream()))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def act_conv1d(x, la, lb, w, b, res, k, d, mode, want_fused=False):
    xt, wt, lat, lbt = dev(x), dev(w), dev(la), dev(lb)
    bt = None if b is None else dev(b)
    rt = None if res is None else dev(res)
    B, Cin, T = xt.shape
    Cout = wt.shape[0]
    y = torch.empty(B, Cout, T, device="cuda")
    fused = C.c_int32(1 if want_fused else 0)   # in: request the fused kernel, out: whether it ran
    lib.check(L_().bvg_act_conv1d(ptr(xt), ptr(lat), ptr(lbt), ptr(wt), ptr(bt), ptr(rt), ptr(y), B, Cin, Cout, T, k, d,
                                  mode, C.addressof(fused), stream()))
    torch.cuda.synchronize()
    return y.cpu().numpy(), bool(fused.value)


def conv_transpose1d(x, w, b, k, u, mode):
    xt, wt = dev(x), dev(w)
    bt = None if b is None else dev(b)
    B, Cin, T = xt.shape
    Cout = wt.shape[1]
    y = torch.empty(B, Cout, T * u, device="cuda")
    lib.check(L_().bvg_conv_transpose1d(ptr(xt), ptr(wt), ptr(bt), ptr(y), B, Cin, Cout, T, k, u, mode, stream()))
    torch.cuda.synchronize()
    return y.cpu().numpy()


def bf16_round(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


def f16_round(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).to(torch.float16).float().numpy().astype(np.float64)
