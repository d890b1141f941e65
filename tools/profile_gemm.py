#!/usr/bin/env python
"""ncu driver: one decode-shape GEMM (M=320, N=1280, K=5120, bias+residual), plain and split-K."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_b200 import _lib, ops
M, N, K = 320, 1280, 5120
x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
b = torch.randn(N, device="cuda").bfloat16(); r = torch.randn(M, N, device="cuda").bfloat16()
scratch = (torch.empty(8 * M * N, device="cuda"), torch.zeros(1024, device="cuda", dtype=torch.int32))
o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
_lib.lib().wb200_set_splitk(1)
for _ in range(3):
    ops.linear(x, w, bias=b, residual=r, out=o); ops.linear_splitk(x, w, bias=b, residual=r, scratch=scratch, out=o)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.linear(x, w, bias=b, residual=r, out=o)
ops.linear_splitk(x, w, bias=b, residual=r, scratch=scratch, out=o)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
