"""Oracle for the model math (reference whisper/model.py).  TEST INFRASTRUCTURE ONLY.

Functional restatement over a flat {name: fp32 tensor} weight dict (the reference's state-dict
names) using torch CPU fp32 tensor arithmetic.  No nn.Module, no hooks: the kv-cache that the
reference builds with forward hooks (model.py:310-341) is an explicit per-layer list here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Weights = Dict[str, torch.Tensor]

# Optional emulation of the reference's 16-bit execution (fp16 on CUDA: every module output is rounded to the
# activation type, matmuls accumulate in fp32, LayerNorm / softmax run in fp32 - model.py:39-50, 124, 247):
# when set, every op output below is rounded to this dtype.  Used to size parity tolerances and decision-margin
# gates, and to pick well-conditioned fixtures; None = pure fp32 (the oracle proper).
ACT_DTYPE: Optional[torch.dtype] = None


def _r(x: torch.Tensor) -> torch.Tensor:
    return x if ACT_DTYPE is None else x.to(ACT_DTYPE).float()


def to_weights(state_dict: Dict[str, np.ndarray], device=None) -> Weights:
    """fp32 tensors, like the reference model after load_model() (parameters stay fp32 on any device; Linear / Conv1d
    cast them to the activation type per call, model.py:44-59)."""
    out = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in state_dict.items()}
    return out if device is None else {k: v.to(device) for k, v in out.items()}


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """model.py:39-41: fp32 LayerNorm over the last dim, eps = nn.LayerNorm default 1e-5."""
    xf = x.float()                                                      # super().forward(x.float()).type(x.dtype)
    mu = xf.mean(-1, keepdim=True)
    var = ((xf - mu) ** 2).mean(-1, keepdim=True)
    return _r((xf - mu) / torch.sqrt(var + 1e-5) * w + b).to(x.dtype)


def gelu(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU: F.gelu / nn.GELU defaults used at model.py:156,193-194."""
    return _r(0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0)))))


def linear(x: torch.Tensor, W: Weights, prefix: str) -> torch.Tensor:
    """model.py:44-50; `key` projections have no bias (model.py:88)."""
    y = x @ W[prefix + ".weight"].to(x.dtype).T                         # self.weight.to(x.dtype), model.py:46-49
    b = W.get(prefix + ".bias")
    return _r(y if b is None else y + b.to(x.dtype))


def conv1d_k3(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    """Conv1d(kernel 3, padding 1) of model.py:178-179 written as three shifted matmuls.
    x: (B, C_in, T) -> (B, C_out, T // stride)."""
    B, C, T = x.shape
    xp = torch.nn.functional.pad(x, (1, 1))
    T_out = (T + 2 - 3) // stride + 1
    w, b = w.to(x.dtype), b.to(x.dtype)                                 # model.py:56-59
    y = torch.zeros(B, w.shape[0], T_out, dtype=x.dtype, device=x.device)
    for k in range(3):
        seg = xp[:, :, k: k + stride * (T_out - 1) + 1: stride]       # (B, C_in, T_out)
        y = y + torch.einsum("oc,bct->bot", w[:, :, k], seg)
    return _r(y + b[None, :, None])


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_head: int,
              causal: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:114-139.  Returns (output, qk) with qk the pre-softmax fp32 scores the non-SDPA
    branch exposes to timing.py.  Scale is d_head^-0.25 on both q and k (== 1/sqrt(d_head))."""
    B, Tq, D = q.shape
    dh = D // n_head
    scale = dh ** -0.25
    qh = q.view(B, Tq, n_head, dh).permute(0, 2, 1, 3)
    kh = k.view(B, k.shape[1], n_head, dh).permute(0, 2, 1, 3)
    vh = v.view(B, v.shape[1], n_head, dh).permute(0, 2, 1, 3)
    qk = (qh * scale) @ (kh * scale).transpose(-1, -2)
    if causal:
        # queries are the LAST Tq positions of the key sequence (kv-cache decoding)
        Tk = kh.shape[2]
        qpos = torch.arange(Tk - Tq, Tk, device=q.device)[:, None]
        kpos = torch.arange(Tk, device=q.device)[None, :]
        qk = qk.masked_fill(kpos > qpos, float("-inf"))
    w = torch.softmax(qk.float(), dim=-1).to(q.dtype)                   # model.py:136
    out = _r((w @ vh).permute(0, 2, 1, 3).reshape(B, Tq, D))
    return out, qk


def encoder_forward(W: Weights, dims: Dict[str, int], mel: torch.Tensor,
                    collect: Optional[dict] = None) -> torch.Tensor:
    """AudioEncoder.forward, model.py:188-204.  mel: (B, n_mels, 3000) -> (B, 1500, d)."""
    x = gelu(conv1d_k3(mel, W["encoder.conv1.weight"], W["encoder.conv1.bias"], 1))
    if collect is not None:
        collect["conv1"] = x
    x = gelu(conv1d_k3(x, W["encoder.conv2.weight"], W["encoder.conv2.bias"], 2))
    x = x.permute(0, 2, 1)
    assert x.shape[1:] == W["encoder.positional_embedding"].shape, "incorrect audio shape"  # model.py:197
    x = _r(x + W["encoder.positional_embedding"]).to(x.dtype)
    if collect is not None:
        collect["stem"] = x
    H = dims["n_audio_head"]
    for i in range(dims["n_audio_layer"]):
        p = f"encoder.blocks.{i}"
        h = layer_norm(x, W[p + ".attn_ln.weight"], W[p + ".attn_ln.bias"])
        a, _ = attention(linear(h, W, p + ".attn.query"), linear(h, W, p + ".attn.key"),
                         linear(h, W, p + ".attn.value"), H, causal=False)
        x = _r(x + linear(a, W, p + ".attn.out"))
        h = layer_norm(x, W[p + ".mlp_ln.weight"], W[p + ".mlp_ln.bias"])
        x = _r(x + linear(gelu(linear(h, W, p + ".mlp.0")), W, p + ".mlp.2"))
        if collect is not None:
            collect[f"block{i}"] = x
    return layer_norm(x, W["encoder.ln_post.weight"], W["encoder.ln_post.bias"])


class KVCache:
    """Explicit form of the reference's hook-built cache (model.py:310-341, decoding.py:144-176):
    per decoder layer, self-attention K/V grown by concatenation and cross-attention K/V computed
    once from the audio features."""

    def __init__(self, n_layer: int):
        self.self_k: List[Optional[torch.Tensor]] = [None] * n_layer
        self.self_v: List[Optional[torch.Tensor]] = [None] * n_layer
        self.cross_k: List[Optional[torch.Tensor]] = [None] * n_layer
        self.cross_v: List[Optional[torch.Tensor]] = [None] * n_layer

    @property
    def length(self) -> int:
        return 0 if self.self_k[0] is None else self.self_k[0].shape[1]

    def reorder(self, source_indices: List[int]) -> None:
        """decoding.py:172-176: only the self-attention caches are gathered."""
        if source_indices != list(range(len(source_indices))):
            idx = torch.tensor(source_indices)
            for i in range(len(self.self_k)):
                self.self_k[i] = self.self_k[i][idx]
                self.self_v[i] = self.self_v[i][idx]


def decoder_forward(W: Weights, dims: Dict[str, int], tokens: torch.Tensor, xa: torch.Tensor,
                    cache: Optional[KVCache] = None, collect_qk: Optional[list] = None) -> torch.Tensor:
    """TextDecoder.forward, model.py:227-249.  tokens: (R, n) int64, xa: (B_or_R, 1500, d);
    returns fp32 logits (R, n, V).  With a cache, `tokens` are the new positions only."""
    offset = cache.length if cache is not None else 0                                  # model.py:234
    x = _r(W["decoder.token_embedding.weight"][tokens] +
           W["decoder.positional_embedding"][offset: offset + tokens.shape[-1]])        # model.py:235-238
    x = x.to(xa.dtype)                                                                 # model.py:239
    H = dims["n_text_head"]
    for i in range(dims["n_text_layer"]):
        p = f"decoder.blocks.{i}"
        h = layer_norm(x, W[p + ".attn_ln.weight"], W[p + ".attn_ln.bias"])
        k_new, v_new = linear(h, W, p + ".attn.key"), linear(h, W, p + ".attn.value")
        if cache is not None:
            if cache.self_k[i] is not None:                                            # model.py:327-333
                k_new = torch.cat([cache.self_k[i], k_new], dim=1)
                v_new = torch.cat([cache.self_v[i], v_new], dim=1)
            cache.self_k[i], cache.self_v[i] = k_new, v_new
        a, _ = attention(linear(h, W, p + ".attn.query"), k_new, v_new, H, causal=True)  # model.py:124-127
        x = _r(x + linear(a, W, p + ".attn.out"))
        h = layer_norm(x, W[p + ".cross_attn_ln.weight"], W[p + ".cross_attn_ln.bias"])
        if cache is not None and cache.cross_k[i] is not None:                          # model.py:106-109
            ck, cv = cache.cross_k[i], cache.cross_v[i]
        else:
            ck, cv = linear(xa, W, p + ".cross_attn.key"), linear(xa, W, p + ".cross_attn.value")
            if cache is not None:
                cache.cross_k[i], cache.cross_v[i] = ck, cv
        a, qk = attention(linear(h, W, p + ".cross_attn.query"), ck, cv, H, causal=False)
        if collect_qk is not None:
            collect_qk.append(qk)
        x = _r(x + linear(a, W, p + ".cross_attn.out"))
        h = layer_norm(x, W[p + ".mlp_ln.weight"], W[p + ".mlp_ln.bias"])
        x = _r(x + linear(gelu(linear(h, W, p + ".mlp.0")), W, p + ".mlp.2"))
    x = layer_norm(x, W["decoder.ln.weight"], W["decoder.ln.bias"])
    return (x @ W["decoder.token_embedding.weight"].to(x.dtype).T).float()             # model.py:245-247
