"""CPU: the multi-GPU plumbing (replicated weights, sharded segments) with world_size 2 over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from whisper_b200 import parallel, synthetic

    r, w, _ = parallel.init_from_env("gloo")
    dims = synthetic.dims_dict("test-en")
    spec = [(n, s) for n, s, _ in synthetic.state_dict_spec(dims)][:12]
    sd = None
    if r == 0:
        full = synthetic.synthetic_state_dict(dims, seed=3)
        sd = {n: full[n] for n, _ in spec}
    got = parallel.broadcast_state_dict(sd, spec, "cpu")
    ref = synthetic.synthetic_state_dict(dims, seed=3)
    ok_bcast = all(np.array_equal(got[n].numpy(), ref[n]) for n, _ in spec)
    lo, hi = parallel.shard_range(7, r, w)
    tokens = [[100 * i + k for k in range(i % 4)] for i in range(lo, hi)]      # includes empty lists
    t, lp, ns = parallel.gather_results(tokens, [float(-i) for i in range(lo, hi)], [i / 10.0 for i in range(lo, hi)], "cpu")
    out[rank] = (ok_bcast, (lo, hi), t, lp, ns)
    import torch.distributed as dist

    dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    port = 29650 + os.getpid() % 200
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert res[0][1] == (0, 4) and res[1][1] == (4, 7)
    expect_tokens = [[100 * i + k for k in range(i % 4)] for i in range(7)]
    for r in (0, 1):
        assert res[r][2] == expect_tokens
        assert res[r][3] == [float(-i) for i in range(7)]
        assert np.allclose(res[r][4], [i / 10.0 for i in range(7)])


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from whisper_b200.parallel import shard_range

    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference algorithm on host cores) prints ONE JSON line with the keys the
    driver reads; run here on the small test model so it takes seconds."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--model", "test-en", "--batch", "2", "--beam", "2"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["impl"] == "reference" and j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["gpu_launches"] == 0
