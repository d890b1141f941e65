"""Word-timing math with the reference's surface (whisper/timing.py:19-151): `median_filter` and
`dtw`, both on the GPU through the C ABI (csrc/timing.cu).  There is no Triton, no numba and no
CPU fallback: the DTW backtrace also runs on the device, only the final path (a few hundred int32)
comes back to the host."""
from __future__ import annotations

from ctypes import c_int, c_int64, c_size_t

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr


def _cuda(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("whisper_b200.timing needs a CUDA device (no CPU path)")
        x = x.cuda()
    return x


def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    """Median filter of width `filter_width` along the last dim with reflect padding (timing.py:19-54)."""
    pad_width = filter_width // 2
    if x.shape[-1] <= pad_width:
        return x                                                    # timing.py:22-24
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    xc = _cuda(x).to(torch.float32).contiguous()
    y = torch.empty_like(xc)
    T = xc.shape[-1]
    rows = xc.numel() // T
    with torch.cuda.device(xc.device):
        check(lib().wb200_median_filter(ptr(xc), ptr(y), c_int64(rows), c_int(T), c_int(filter_width), stream_ptr()),
              "wb200_median_filter")
    return y.to(x.dtype) if x.dtype != torch.float32 else y


def dtw(x: torch.Tensor, cpu_tie_break: bool = False) -> np.ndarray:
    """Dynamic time warping over cost matrix x (N text tokens, M frames) -> int array (2, path_len)
    (timing.py:82-151).  Ties follow the rule the reference applies to CUDA tensors
    (triton_ops.py:38-40) unless `cpu_tie_break` asks for dtw_cpu's (timing.py:95-100)."""
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x))
    xc = _cuda(x).to(torch.float32).contiguous()
    N, M = xc.shape
    path = torch.empty((2, N + M + 1), device=xc.device, dtype=torch.int32)
    n = torch.zeros(1, device=xc.device, dtype=torch.int32)
    with torch.cuda.device(xc.device):
        nbytes = int(lib().wb200_dtw_workspace_bytes(c_int(N), c_int(M)))
        ws = torch.empty(nbytes, device=xc.device, dtype=torch.uint8)
        check(lib().wb200_dtw(ptr(xc), c_int(N), c_int(M), ptr(path), ptr(n), ptr(ws), c_size_t(nbytes),
                              c_int(int(cpu_tie_break)), stream_ptr()), "wb200_dtw")
    length = int(n.item())
    return path[:, :length].cpu().numpy().astype(np.int64)
