"""CPU: the reference's own known-answer tests for the word-timing math, run against the oracle
(reference tests/test_timing.py:22-52 planted DTW path, :67-84 median filter vs scipy)."""
import numpy as np
import pytest
import scipy.ndimage

from oracle import timing as OT

sizes = [(10, 20), (32, 16), (123, 1500), (234, 189)]
shapes = [(10,), (1, 15), (4, 5, 345), (6, 12, 240, 512)]


@pytest.mark.parametrize("N, M", sizes)
def test_dtw_planted_path(N, M):
    rng = np.random.RandomState(42)
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
    trace = np.array(trace).T
    assert np.array_equal(trace, OT.dtw(x))
    assert np.array_equal(trace, OT.dtw_gpu_tiebreak(x))      # no exact ties on random input


@pytest.mark.parametrize("shape", shapes)
def test_median_filter_vs_scipy(shape):
    rng = np.random.RandomState(0)
    x = rng.randn(*shape).astype(np.float32)
    for width in [3, 5, 7, 13]:
        got = OT.median_filter(x, width)
        pad = width // 2
        padded = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
        ref = scipy.ndimage.median_filter(padded, [1] * (x.ndim - 1) + [width])[..., pad:-pad]
        assert np.array_equal(got, ref)


def test_median_filter_short_input_passthrough():
    x = np.arange(3, dtype=np.float32)
    assert OT.median_filter(x, 7) is x                       # timing.py:22-24
