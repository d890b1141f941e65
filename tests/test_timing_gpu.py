"""GPU: median filter and DTW kernels - bit-exact (comparison-only / integer path outputs)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD

pytestmark = pytest.mark.gpu


def test_median_filter_golden_and_oracle():
    from oracle import timing as OT
    from whisper_b200.timing import median_filter

    g = np.load(os.path.join(GOLD, "timing.npz"))
    for i in range(4):
        x = g[f"med_in_{i}"]
        for w in (3, 5, 7, 13):
            got = median_filter(torch.from_numpy(x).cuda(), w).cpu().numpy()
            assert np.array_equal(got, g[f"med_out_{i}_{w}"]), (i, w)
    rng = np.random.RandomState(1)
    for shape in [(10,), (1, 15), (4, 5, 345), (6, 12, 240, 512), (23, 448, 1500)]:   # reference test shapes + C3-like
        x = rng.randn(*shape).astype(np.float32)
        for w in (3, 5, 7, 13):
            got = median_filter(torch.from_numpy(x).cuda(), w).cpu().numpy()
            assert np.array_equal(got, OT.median_filter(x, w)), (shape, w)
    short = torch.arange(3.0, device="cuda")
    assert median_filter(short, 7) is short                       # timing.py:22-24


@pytest.mark.parametrize("N,M", [(10, 20), (32, 16), (123, 1500), (234, 189), (1, 1), (1, 40), (40, 1), (448, 1500)])
def test_dtw_random(N, M):
    from oracle import timing as OT
    from whisper_b200.timing import dtw

    rng = np.random.RandomState(N * 1000 + M)
    x = rng.randn(N, M).astype(np.float32)
    ref = OT.dtw(x)
    assert np.array_equal(dtw(torch.from_numpy(x).cuda()), ref)
    assert np.array_equal(dtw(torch.from_numpy(x).cuda(), cpu_tie_break=True), ref)


def test_dtw_golden():
    from whisper_b200.timing import dtw

    g = np.load(os.path.join(GOLD, "timing.npz"))
    for i in range(4):
        assert np.array_equal(dtw(torch.from_numpy(g[f"dtw_in_{i}"]).cuda()), g[f"dtw_out_{i}"])


def test_dtw_exact_ties_follow_both_reference_rules():
    """Integer-valued costs make ties common: the CUDA rule (triton_ops.py:38-40) and the CPU rule
    (timing.py:95-100) give different paths; each mode must match its oracle exactly."""
    from oracle import timing as OT
    from whisper_b200.timing import dtw

    rng = np.random.RandomState(5)
    for N, M in [(7, 9), (30, 50), (64, 200)]:
        x = rng.randint(0, 3, size=(N, M)).astype(np.float32)
        assert np.array_equal(dtw(torch.from_numpy(x).cuda()), OT.dtw_gpu_tiebreak(x))
        assert np.array_equal(dtw(torch.from_numpy(x).cuda(), cpu_tie_break=True), OT.dtw(x))


def test_dtw_planted_path():
    """Reference tests/test_timing.py:22-52."""
    from whisper_b200.timing import dtw

    rng = np.random.RandomState(42)
    for N, M in [(10, 20), (32, 16), (123, 1500), (234, 189)]:
        steps = np.concatenate([np.zeros(N - 1), np.ones(M - 1)])
        rng.shuffle(steps)
        x = rng.random((N, M)).astype(np.float32)
        i, j, k = 0, 0, 0
        trace = []
        while True:
            x[i, j] -= 1
            trace.append((i, j))
            if k == len(steps):
                break
            if k + 1 < len(steps) and steps[k] != steps[k + 1]:
                i, j, k = i + 1, j + 1, k + 2
                continue
            if steps[k] == 0:
                i += 1
            if steps[k] == 1:
                j += 1
            k += 1
        assert np.array_equal(np.array(trace).T, dtw(torch.from_numpy(x).cuda()))
