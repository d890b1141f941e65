"""GPU parity tests of the primitive kernels against plain PyTorch fp32 math.

Floating-point tolerance: every kernel accumulates in fp32 and rounds once to the 16-bit output
type, so results must agree with an fp32 reference (rounded at the same points) to within one
16-bit ulp: rtol 2^-7 for bf16, 2^-10 for fp16, plus a small absolute term for cancellation.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16]


def tol(dtype):
    return (2.0 ** -7, 2e-2) if dtype == torch.bfloat16 else (2.0 ** -10, 3e-3)


def report(name, got, ref, dtype, scale=1.0):
    got = got.float()
    ref = ref.float()
    rtol, atol = tol(dtype)
    atol *= scale
    err = (got - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    nbad = int(bad.sum())
    msg = (f"{name}: max_abs_err={float(err.max()):.4g} ref_absmax={float(ref.abs().max()):.4g} "
           f"bad={nbad}/{ref.numel()}")
    if nbad:
        idx = bad.nonzero()[:5].tolist()
        msg += f" first_bad={idx} got={[float(got[tuple(i)]) for i in idx]} ref={[float(ref[tuple(i)]) for i in idx]}"
    assert nbad == 0, msg


def rt(x, dtype):
    return x.to(dtype).float()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (77, 64, 128), (300, 1280, 1280), (320, 1536, 384),
                                   (4500, 1280, 1280), (4096, 512, 2048), (5000, 384, 1536)])
def test_linear_plain(dtype, M, N, K):
    from whisper_b200 import ops
    torch.manual_seed(0)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda") * (1.0 / math.sqrt(K))).to(dtype)
    y = ops.linear(x, w)
    ref = rt(x.float() @ w.float().T, dtype)
    report(f"linear {M}x{N}x{K}", y, ref, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M", [320, 4500])
def test_linear_epilogues(dtype, M):
    from whisper_b200 import ops
    torch.manual_seed(1)
    N, K = 1280, 512
    x = (torch.randn(M, K, device="cuda")).to(dtype)
    w = (torch.randn(N, K, device="cuda") * (1.0 / math.sqrt(K))).to(dtype)
    b = (torch.randn(N, device="cuda") * 0.3).to(dtype)
    r = torch.randn(M, N, device="cuda").to(dtype)
    acc = x.float() @ w.float().T
    # bias
    report("bias", ops.linear(x, w, bias=b), rt(acc + b.float(), dtype), dtype)
    # bias + gelu (exact erf)
    g = torch.nn.functional.gelu(rt(acc + b.float(), dtype))
    report("bias+gelu", ops.linear(x, w, bias=b, gelu=True), rt(g, dtype), dtype)
    # bias + residual (two roundings, like Linear output then x + y in the 16-bit type)
    rr = rt(rt(acc + b.float(), dtype) + r.float(), dtype)
    report("bias+residual", ops.linear(x, w, bias=b, residual=r), rr, dtype)
    # residual aliasing the output (in-place x += f(x))
    r2 = r.clone()
    ops.linear(x, w, bias=b, residual=r2, out=r2)
    report("in-place residual", r2, rr, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(320, 1280, 1280), (320, 1280, 5120), (320, 3840, 1280), (5, 1280, 5120),
                                   (64, 384, 1536), (200, 512, 2048)])
def test_linear_splitk_decode_shapes(dtype, M, N, K, monkeypatch):
    """The skinny decode-step GEMMs (C3: 320 rows) through the split-K path (opt-in, WB200_SPLITK=1): same
    result as the plain kernel up to fp32 summation order, epilogues included, tickets back to zero."""
    from whisper_b200 import _lib, ops
    _lib.lib().wb200_set_splitk(1)
    monkeypatch.setattr(ops, "_splitk_restore", True, raising=False)
    torch.manual_seed(7)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda") * (1.0 / math.sqrt(K))).to(dtype)
    b = (torch.randn(N, device="cuda") * 0.3).to(dtype)
    r = torch.randn(M, N, device="cuda").to(dtype)
    acc = x.float() @ w.float().T
    report("splitk plain", ops.linear_splitk(x, w), rt(acc, dtype), dtype)
    report("splitk bias+gelu", ops.linear_splitk(x, w, bias=b, gelu=True),
           rt(torch.nn.functional.gelu(rt(acc + b.float(), dtype)), dtype), dtype)
    report("splitk bias+residual", ops.linear_splitk(x, w, bias=b, residual=r),
           rt(rt(acc + b.float(), dtype) + r.float(), dtype), dtype)
    got = ops.linear_splitk(x, w, out_f32=True)
    _lib.lib().wb200_set_splitk(0)
    assert float((got - acc).abs().max()) < 2e-3 * math.sqrt(K) / 30 + 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
def test_linear_f32_out_ragged_n(dtype):
    """Logits-shaped product: N not a multiple of the tile, fp32 output with a padded row stride."""
    from whisper_b200 import ops
    torch.manual_seed(2)
    M, N, K = 35, 5187, 384
    x = torch.randn(M, K, device="cuda").to(dtype)
    w = torch.randn(N, K, device="cuda").to(dtype)
    ld = (N + 31) // 32 * 32
    buf = torch.full((M, ld), 777.0, device="cuda", dtype=torch.float32)
    out = buf[:, :N]
    ops.linear(x, w, out_f32=True, out=out)
    ref = x.float() @ w.float().T
    err = (out - ref).abs().max()
    assert float(err) < 2e-2 * math.sqrt(K) / 20, f"fp32-out max err {float(err)}"
    assert bool((buf[:, N:] == 777.0).all()), "padding columns were overwritten"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,T,Cin,Cout,stride", [(2, 3000, 80, 384, 1), (3, 3000, 128, 256, 1),
                                                 (2, 3000, 384, 384, 2), (2, 200, 128, 128, 2),
                                                 (1, 130, 80, 64, 1)])
def test_conv1d_k3_gelu(dtype, B, T, Cin, Cout, stride):
    from whisper_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(B, Cin, T, device="cuda").to(dtype)           # reference layout (B, C, T)
    w = (torch.randn(Cout, Cin, 3, device="cuda") / math.sqrt(3 * Cin)).to(dtype)
    b = (torch.randn(Cout, device="cuda") * 0.2).to(dtype)
    pos = torch.randn(T // stride, Cout, device="cuda") if stride == 2 else None
    xt = x.transpose(1, 2).contiguous()                           # time-major
    wt = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()   # tap-major
    y = ops.conv1d_k3_gelu(xt, wt, b, stride=stride, pos=pos)
    conv = torch.nn.functional.conv1d(x.float(), w.float(), b.float(), stride=stride, padding=1)
    ref = rt(torch.nn.functional.gelu(rt(conv, dtype)), dtype).transpose(1, 2)
    if pos is not None:
        ref = rt(ref + pos, dtype)
    # the positional add can cancel: a 1-ulp flip of the (larger) GELU output survives it
    report(f"conv s{stride}", y, ref, dtype, scale=4.0 if pos is not None else 1.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,d", [(1, 384), (37, 512), (3000, 1280), (5, 2048)])
def test_layernorm(dtype, rows, d):
    from whisper_b200 import ops
    torch.manual_seed(4)
    x = (torch.randn(rows, d, device="cuda") * 3 + 0.5).to(dtype)
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    y = ops.layernorm(x, g, b)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-5)
    report("layernorm", y, rt(ref, dtype), dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_transpose_to16(dtype):
    from whisper_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(3, 80, 3000, device="cuda")
    y = ops.transpose_to16(x, dtype)
    assert torch.equal(y, x.transpose(1, 2).to(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,T,H", [(2, 1500, 6), (1, 300, 2), (3, 256, 1), (1, 1, 1), (2, 129, 3)])
def test_encoder_attention(dtype, B, T, H):
    from whisper_b200 import ops
    torch.manual_seed(6)
    d = 64 * H
    qkv = (torch.randn(B * T, 3 * d, device="cuda") * 1.5).to(dtype)
    out = ops.encoder_attention(qkv, B, T, H)
    q, k, v = qkv.float().reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = s.softmax(-1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * T, d)
    # P is rounded to 16 bits before the PV product (as in SDPA's flash kernels): widen atol
    report("enc attention", out, rt(ref, dtype), dtype, scale=1.5)
