"""CPU oracle for the Whisper inference hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithms of the reference
(openai/whisper @ c0d2f62) for the one path whisper_b200 accelerates:

    log_mel_spectrogram -> AudioEncoder -> TextDecoder (kv-cache) -> logit filters ->
    greedy / beam-search selection -> ranking, plus timing.py's median filter and DTW.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it; the
product package `whisper_b200` never does (tests/test_no_oracle_in_product.py enforces that).

Parity pinning: the reference is a Python package and imports in the build container, so every
function here is checked against the reference itself (tests/test_oracle_vs_reference.py, which
runs whenever /root/reference exists) and against golden vectors generated from the reference by
oracle/make_golden.py and committed under tests/golden/ (those travel to the GPU box, where the
reference does not exist).  The reference's own known-answer tests for this path
(tests/test_timing.py: planted DTW path, scipy median filter) are reproduced in
tests/test_oracle_timing.py.

Numerics: neural-network math uses torch CPU fp32 tensors (the reference's own CPU arithmetic:
F.linear / softmax / erf-GELU in fp32); token-selection logic is plain Python / numpy on integers
and fp32 scalars; the mel front-end uses numpy float64 internally and rounds to fp32 once.
"""
