"""Step-wise parity drivers shared by tests/, __graft_entry__.smoke() and bench.py's parity check.
TEST INFRASTRUCTURE ONLY (like the rest of oracle/): the product never imports this.

Two ways of comparing a device decoder session (whisper_b200.decoding.DecoderSession, i.e. the C ABI) with the
fp32 oracle, both following the reference's loop (whisper/decoding.py:680-710):

* teacher_forced(): the oracle's fp32 logits are injected into the device selection kernels at every step, so the
  device follows the oracle's trajectory exactly - including every beam reorder (decoding.py:172-176,
  `rearrange_kv_cache`) - and the logits the DEVICE computed for that step (its kv-cache read through the
  parent table after real reorders) are compared with the oracle's before they are overwritten.
* free_running(): the device decodes on its own logits.  A 16-bit pipeline cannot reproduce decisions whose margin
  is below its rounding noise, so every step is gated on a MEASURED bound: the step's decisions must be identical
  whenever the oracle's smallest decision gap exceeds 2 x (error of the accumulated scores + error of this
  step's candidate log-probabilities), both measured against the oracle on the spot.  The number of steps
  asserted is returned.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import decoding as OD


def oracle_record(W, dims, feats, opts: dict, n_audio: int, max_steps: Optional[int] = None) -> dict:
    """Run the oracle on `n_audio` feature rows and return its per-step record (raw / filtered logits, tokens,
    beam parents, score sums, decision gaps)."""
    o = dict(opts)
    if o.get("suppress_tokens") == "":
        o["suppress_tokens"] = ()
    oopt = OD.Options(**o)
    rec: dict = {}
    res = OD.decode(W, dims, feats[:n_audio], oopt, record=rec, max_steps=max_steps)
    rec["results"] = res
    rec["options"] = oopt
    return rec


def open_session(model, opts: dict, n_audio: int, g_feats):
    from whisper_b200.decoding import DecodingOptions, DecodingTask

    task = DecodingTask(model, DecodingOptions(language=opts.get("language", "en"),
                                               **{k: v for k, v in opts.items() if k != "language"}))
    sess = task.open_session(n_audio)
    sess.set_audio(g_feats[:n_audio].contiguous())
    sess.prefill(np.tile(np.asarray(task.initial_tokens, dtype=np.int32), (n_audio, 1)))
    return task, sess


def teacher_forced(model, opts: dict, n_audio: int, g_feats, rec: dict, logit_tol: float) -> Dict[str, float]:
    """Inject the oracle's logits step by step; compare the device's own logits of every step with the oracle's
    (relative to the largest |logit| of the step) and require identical tokens / parents / score sums."""
    task, sess = open_session(model, opts, n_audio, g_feats)
    G = task.n_group
    R = n_audio * G
    worst, reorders = 0.0, 0
    try:
        n_steps = len(rec["raw_logits"])
        for i in range(n_steps):
            ref = rec["raw_logits"][i]
            if i == 0:
                got = sess.get_logits(n_audio).float().cpu()
                cmp_ref = ref[::G]
            else:
                sess.step()
                got = sess.get_logits(R).float().cpu()
                cmp_ref = ref
            assert bool(torch.isfinite(got).all()), f"step {i}: non-finite device logits"
            err = float((got - cmp_ref).abs().max() / cmp_ref.abs().max())
            worst = max(worst, err)
            assert err < logit_tol, f"step {i}: device logits deviate from the oracle by {err:.5f} of max|logit| (tol {logit_tol})"
            sess.set_logits(ref[::G] if i == 0 else ref)
            sess.select()
            L = int(sess.get("length").item())
            toks = sess.get("tokens")[:, :L].cpu().numpy().tolist()
            assert toks == rec["tokens_out"][i], f"step {i}: tokens differ under injected logits"
            lp = sess.get("sum_logprobs").cpu()
            ref_lp = rec["sum_logprobs_out"][i]
            live = torch.isfinite(ref_lp)
            assert torch.allclose(lp[live], ref_lp[live], atol=1e-4, rtol=1e-5), f"step {i}: sum_logprobs differ"
            if G > 1 and "source_indices" in rec:
                src = sess.get("sources").cpu().tolist()
                assert src == rec["source_indices"][i], f"step {i}: beam parents differ"
                reorders += int(src != list(range(R)))
        done = int(sess.get("done").item())
    finally:
        sess.close()
    return {"worst_rel_logit_err": worst, "steps": n_steps, "reorders": reorders, "done": done}


def free_running(model, opts: dict, n_audio: int, g_feats, rec: dict, dims: Dict[str, int]) -> Dict[str, float]:
    """Let the device decode on its own logits and assert its decisions step by step while the measured error bound
    stays below the oracle's decision gaps (see the module docstring)."""
    task, sess = open_session(model, opts, n_audio, g_feats)
    G = task.n_group
    R = n_audio * G
    oopt: OD.Options = rec["options"]
    beam = oopt.beam_size is not None
    ids, _init, sample_begin, sup, mits = OD.filter_context(dims, oopt)
    asserted, E, first_gap, first_bound = 0, 0.0, None, None
    try:
        n_steps = len(rec["raw_logits"])
        for i in range(n_steps):
            if i > 0:
                sess.step()
            got = sess.get_logits(n_audio if i == 0 else R).float().cpu()
            if i == 0 and G > 1:
                got = got.repeat_interleave(G, dim=0)
            dev = got.clone()
            OD.apply_filters(dev, rec["tokens_in"][i], ids, oopt, sample_begin, sup, mits)
            lp_dev = torch.log_softmax(dev, dim=-1)
            lp_ora = torch.log_softmax(rec["filtered_logits"][i].float(), dim=-1)
            sess.select()
            K = G + 1 if beam else 1
            cand = lp_ora.topk(min(K + 1, lp_ora.shape[-1]), dim=-1).indices
            dev_top = sess.get("top_idx").cpu().long().reshape(R, -1)
            cand = torch.cat([cand, dev_top], dim=-1)
            a, b = lp_dev.gather(-1, cand), lp_ora.gather(-1, cand)
            both_inf = torch.isinf(a) & torch.isinf(b) & (a == b)
            diff = torch.where(both_inf, torch.zeros_like(a), (a - b).abs())
            e = float(torch.nan_to_num(diff, nan=float("inf")).max())
            if beam:
                gap = rec["beam_step_gaps"][i]
            else:
                gap = min(rec["all_margins"][r][i] for r in range(R))
            bound = 2.0 * (E + e) if beam else 2.0 * e      # greedy decisions do not depend on the running sums
            if first_gap is None:
                first_gap, first_bound = gap, bound
            if not (gap > bound):
                break
            L = int(sess.get("length").item())
            toks = sess.get("tokens")[:, :L].cpu().numpy().tolist()
            assert toks == rec["tokens_out"][i], (f"free-running step {i}: tokens differ although the decision gap "
                                                  f"{gap:.4f} exceeds the measured error bound {bound:.4f}")
            if beam:
                src = sess.get("sources").cpu().tolist()
                assert src == rec["source_indices"][i], f"free-running step {i}: beam parents differ (gap {gap:.4f}, bound {bound:.4f})"
            lp = sess.get("sum_logprobs").cpu()
            ref_lp = rec["sum_logprobs_out"][i]
            live = torch.isfinite(ref_lp)
            E = float((lp[live] - ref_lp[live]).abs().max()) if bool(live.any()) else E
            asserted += 1
    finally:
        sess.close()
    return {"asserted_steps": asserted, "steps": n_steps, "first_gap": first_gap, "first_bound": first_bound}
