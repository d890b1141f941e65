"""Multi-GPU plumbing: one process per GPU, replicated weights, segments sharded across ranks.

`decode()` over a batch of 30-second segments shards naturally: no kv-cache, beam or
filter state crosses segments, so there is NO per-step collective.  torch.distributed (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests) is used exactly twice per job:

  * `broadcast_state_dict`  - rank 0 owns the checkpoint; everyone else receives it (one flat buffer
                              per dtype, so the 3 GB of large-v3 moves in two collectives);
  * `gather_results`        - variable-length token lists, log-probs and no-speech probs of every
                              rank's segments come back in rank order.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group
    when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items owned by `rank` (rank r gets [r*B/W, (r+1)*B/W));
    remainders go to the lowest ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state_dict(state_dict: Optional[Dict[str, np.ndarray]], spec: Sequence[Tuple[str, Tuple[int, ...]]],
                         device, src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` passes the fp32 state dict; every rank returns {name: fp32 tensor on `device`}.
    `spec` = [(name, shape)] must be identical on all ranks (whisper_b200.synthetic.state_dict_spec)."""
    total = sum(int(np.prod(shape)) for _, shape in spec)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if not dist.is_initialized() or dist.get_world_size() == 1 or dist.get_rank() == src:
        assert state_dict is not None
        off = 0
        for name, shape in spec:
            n = int(np.prod(shape))
            v = state_dict[name]
            v = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
            flat[off: off + n].copy_(v.reshape(-1).to(torch.float32))
            off += n
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    out, off = {}, 0
    for name, shape in spec:
        n = int(np.prod(shape))
        out[name] = flat[off: off + n].view(*shape)
        off += n
    return out


def gather_results(tokens: List[List[int]], avg_logprobs: List[float], no_speech: List[float], device,
                   max_len: int = 448):
    """All ranks contribute their segments' results; returns (tokens, avg_logprobs, no_speech_probs)
    for ALL segments in rank order on every rank.  Fixed-size padded buffers keep it to one
    all_gather per field; ranks may own different numbers of segments."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tokens, avg_logprobs, no_speech
    world = dist.get_world_size()
    n_local = torch.tensor([len(tokens)], device=device, dtype=torch.int64)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    tok = torch.full((n_max, max_len + 1), -1, device=device, dtype=torch.int32)
    meta = torch.zeros((n_max, 2), device=device, dtype=torch.float32)
    for i, t in enumerate(tokens):
        tok[i, 0] = len(t)
        if len(t):
            tok[i, 1: 1 + len(t)] = torch.tensor(t, dtype=torch.int32, device=device)
        meta[i, 0] = avg_logprobs[i]
        meta[i, 1] = no_speech[i]
    toks = [torch.empty_like(tok) for _ in range(world)]
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(toks, tok)
    dist.all_gather(metas, meta)
    out_t, out_lp, out_ns = [], [], []
    for r in range(world):
        tt, mm = toks[r].cpu(), metas[r].cpu()
        for i in range(counts[r]):
            n = int(tt[i, 0])
            out_t.append(tt[i, 1: 1 + n].tolist())
            out_lp.append(float(mm[i, 0]))
            out_ns.append(float(mm[i, 1]))
    return out_t, out_lp, out_ns
