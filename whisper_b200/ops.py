"""Python-side wrappers of the primitive C-ABI operators (used by the host model and the tests).

Every function launches hand-written sm_100a kernels through libwhisper_b200.so on torch's current
CUDA stream; torch only provides the device memory.
"""
from __future__ import annotations

from ctypes import c_int, c_int64

import torch

from ._lib import check, dtype_code, lib, ptr, stream_ptr


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("whisper_b200 operators require CUDA tensors (there is no CPU path)")


def linear(x, weight, bias=None, residual=None, gelu=False, out_f32=False, out=None):
    """y = x @ weight.T (+bias) (GELU) (+residual); reference whisper/model.py:44-50."""
    _req_cuda(x, weight, bias, residual)
    assert x.dtype == weight.dtype and x.dim() == 2 and weight.dim() == 2
    M, K = x.shape
    N = weight.shape[0]
    assert weight.shape[1] == K and x.stride(1) == 1 and weight.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    check(lib().wb200_linear(
        c_int(dtype_code(x.dtype)), c_int(M), c_int(N), c_int(K), ptr(x), c_int64(x.stride(0)),
        ptr(weight), c_int64(weight.stride(0)), ptr(bias), ptr(residual),
        c_int64(residual.stride(0) if residual is not None else 0), ptr(out), c_int64(out.stride(0)),
        c_int(int(gelu)), c_int(int(out_f32)), stream_ptr()), "wb200_linear")
    return out


def linear_splitk(x, weight, bias=None, residual=None, gelu=False, out_f32=False, scratch=None, out=None):
    """linear() with the split-K scratch the decoder session uses for its skinny GEMMs.
    scratch = (fp32 workspace, zeroed int32 tickets) may be passed in to keep allocations out of a timed loop."""
    from ctypes import c_size_t

    _req_cuda(x, weight, bias, residual)
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    check_tickets = scratch is None
    if scratch is None:
        scratch = (torch.empty(8 * M * N, device=x.device, dtype=torch.float32),
                   torch.zeros(1024, device=x.device, dtype=torch.int32))
    ws, tickets = scratch
    check(lib().wb200_linear_splitk(
        c_int(dtype_code(x.dtype)), c_int(M), c_int(N), c_int(K), ptr(x), c_int64(x.stride(0)),
        ptr(weight), c_int64(weight.stride(0)), ptr(bias), ptr(residual),
        c_int64(residual.stride(0) if residual is not None else 0), ptr(out), c_int64(out.stride(0)),
        c_int(int(gelu)), c_int(int(out_f32)), ptr(ws), c_size_t(ws.numel() * 4), ptr(tickets), c_int(1024),
        stream_ptr()), "wb200_linear_splitk")
    if check_tickets:
        assert int(tickets.abs().sum()) == 0, "split-K tickets were not reset"
    return out


def conv1d_k3_gelu(x, w_tapmajor, bias, stride=1, pos=None):
    """GELU(Conv1d(k=3, pad=1, stride)) on time-major x [B, T, C_in]; reference model.py:193-194."""
    _req_cuda(x, w_tapmajor, bias, pos)
    B, T, C_in = x.shape
    C_out = w_tapmajor.shape[0]
    assert w_tapmajor.shape[1] == 3 * C_in and x.is_contiguous() and w_tapmajor.is_contiguous()
    y = torch.empty((B, T // stride, C_out), device=x.device, dtype=x.dtype)
    check(lib().wb200_conv1d_k3_gelu(
        c_int(dtype_code(x.dtype)), c_int(B), c_int(T), c_int(C_in), c_int(C_out), c_int(stride),
        ptr(x), ptr(w_tapmajor), ptr(bias), ptr(pos), ptr(y), stream_ptr()), "wb200_conv1d_k3_gelu")
    return y


def layernorm(x, gamma, beta):
    """LayerNorm in fp32 over the last dim; reference model.py:39-41."""
    _req_cuda(x, gamma, beta)
    d = x.shape[-1]
    x2 = x.reshape(-1, d)
    assert x2.is_contiguous() and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    y = torch.empty_like(x2)
    check(lib().wb200_layernorm(c_int(dtype_code(x.dtype)), ptr(x2), ptr(y), ptr(gamma), ptr(beta),
                                c_int(x2.shape[0]), c_int(d), stream_ptr()), "wb200_layernorm")
    return y.reshape(x.shape)


def transpose_to16(x, dtype):
    """(B, C, T) fp32 -> (B, T, C) 16-bit."""
    _req_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, C, T = x.shape
    y = torch.empty((B, T, C), device=x.device, dtype=dtype)
    check(lib().wb200_transpose_to16(c_int(dtype_code(dtype)), ptr(x), ptr(y), c_int(B), c_int(C),
                                     c_int(T), stream_ptr()), "wb200_transpose_to16")
    return y


def encoder_attention(qkv, B, T, n_head):
    """softmax(q k^T / 8) v per head on packed qkv [B*T, 3*d]; reference model.py:114-139."""
    _req_cuda(qkv)
    d = n_head * 64
    assert qkv.shape == (B * T, 3 * d) and qkv.is_contiguous()
    out = torch.empty((B * T, d), device=qkv.device, dtype=qkv.dtype)
    check(lib().wb200_encoder_attention(c_int(dtype_code(qkv.dtype)), ptr(qkv), ptr(out), c_int(B),
                                        c_int(T), c_int(n_head), stream_ptr()), "wb200_encoder_attention")
    return out
