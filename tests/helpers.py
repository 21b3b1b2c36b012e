"""Shared helpers for the GPU parity tests (checker side only)."""
import ctypes as C

import torch

from teochat_b200 import lib as L


def stream():
    return torch.cuda.current_stream().cuda_stream


def bf(x):
    return x.to(torch.bfloat16)


def rnd(*shape, scale=1.0, seed=0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device)


def gemm(teo, A, W, bias=None, residual=None, act=0, out_fp32=False, C_out=None):
    lib, h = teo
    M, K = A.shape
    N = W.shape[0]
    out = C_out if C_out is not None else torch.empty(M, N, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=A.device)
    wsb = lib.teo_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=A.device)
    L.check(lib.teo_gemm_bf16(h, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), out.data_ptr(), out.stride(0), M, N, K,
                              L.ptr(bias), L.ptr(residual), residual.stride(0) if residual is not None else 0, act,
                              1 if out_fp32 else 0, ws.data_ptr(), ws.numel(), stream()), "teo_gemm_bf16")
    return out


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
