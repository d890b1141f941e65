#!/usr/bin/env python
"""Micro-benchmark of the decode-step GEMM shapes: plain vs split-K tcgen05 kernel (CUDA events, L2 flushed)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
for (M, N, K) in [(320, 1280, 1280), (320, 3840, 1280), (320, 5120, 1280), (320, 1280, 5120), (5, 1280, 1280), (5, 1280, 5120)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); r = torch.randn(M, N, device="cuda").bfloat16()
    scratch = (torch.empty(8 * M * N, device="cuda"), torch.zeros(1024, device="cuda", dtype=torch.int32))
    o0 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); o1 = torch.empty_like(o0)
    t0 = timeit(lambda: ops.linear(x, w, bias=b, residual=r, out=o0))
    t1 = timeit(lambda: ops.linear_splitk(x, w, bias=b, residual=r, scratch=scratch, out=o1))
    t2 = timeit(lambda: torch.nn.functional.linear(x, w, b))
    gb = N * K * 2 / 1e9
    print(f"M={M} N={N} K={K}: plain {t0:7.1f} us  splitk {t1:7.1f} us  (cuBLAS via torch {t2:7.1f} us)  weights {gb*1e3:.1f} MB -> floor {gb/6.5*1e3:.1f} us")
