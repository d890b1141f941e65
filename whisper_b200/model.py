"""`Whisper` model object with the reference's attribute surface (whisper/model.py:252-345), backed
by the C-ABI library: the weights live in caller-owned torch tensors (converted once to the 16-bit
compute type and re-laid-out for the kernels), every forward pass is a call into libwhisper_b200.so.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_size_t, c_void_p
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from ._lib import WhisperB200Error, check, dtype_code, lib, ptr, stream_ptr


@dataclass
class ModelDimensions:
    """Reference whisper/model.py:25-36."""
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int

    def as_list(self):
        return [self.n_mels, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer,
                self.n_vocab, self.n_text_ctx, self.n_text_state, self.n_text_head, self.n_text_layer]


def _t(sd, name, device, dtype):
    v = sd[name]
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v))
    return v.to(device=device, dtype=dtype).contiguous()


def pack_weights(sd: Dict[str, "np.ndarray | torch.Tensor"], dims: ModelDimensions, device, dtype):
    """State dict (reference names) -> flat tensor list in the slot order of include/whisper_b200.h."""
    T, F32 = dtype, torch.float32
    d = dims.n_audio_state
    out = []

    def conv_w(name):   # [out, in, 3] -> tap-major [out, 3*in]
        w = _t(sd, name, device, F32)
        return w.permute(0, 2, 1).reshape(w.shape[0], -1).to(T).contiguous()

    out += [conv_w("encoder.conv1.weight"), _t(sd, "encoder.conv1.bias", device, T),
            conv_w("encoder.conv2.weight"), _t(sd, "encoder.conv2.bias", device, T),
            _t(sd, "encoder.positional_embedding", device, F32),
            _t(sd, "encoder.ln_post.weight", device, F32), _t(sd, "encoder.ln_post.bias", device, F32),
            _t(sd, "decoder.token_embedding.weight", device, T),
            _t(sd, "decoder.token_embedding.weight", device, F32),
            _t(sd, "decoder.positional_embedding", device, F32),
            _t(sd, "decoder.ln.weight", device, F32), _t(sd, "decoder.ln.bias", device, F32)]

    def fused(prefix, names, with_bias):
        w = torch.cat([_t(sd, f"{prefix}.{n}.weight", device, T) for n in names], 0).contiguous()
        bs = []
        for n, has in zip(names, with_bias):
            bs.append(_t(sd, f"{prefix}.{n}.bias", device, T) if has else torch.zeros(d, device=device, dtype=T))
        return w, torch.cat(bs).contiguous()

    def lin(prefix):
        return [_t(sd, prefix + ".weight", device, T), _t(sd, prefix + ".bias", device, T)]

    def ln(prefix):
        return [_t(sd, prefix + ".weight", device, F32), _t(sd, prefix + ".bias", device, F32)]

    def folded(w, b, ln_prefix):
        """LayerNorm folded into the Linear that consumes it (csrc/dec_layer.cu): with y = LN(x) W^T + b,
        LN(x) = (x - mean) * rstd * gamma + beta, the kernel multiplies RAW rows by wf = W * gamma and finishes with
        y = rstd * (x wf^T - mean * c1) + c2.  c1 sums wf AS STORED in the 16-bit type so that the mean term cancels
        exactly what the tensor cores accumulate."""
        g, beta = _t(sd, ln_prefix + ".weight", device, F32), _t(sd, ln_prefix + ".bias", device, F32)
        wf = (w.to(F32) * g[None, :]).to(T).contiguous()
        c1 = wf.to(F32).sum(dim=1).contiguous()
        c2 = (w.to(F32) @ beta + b.to(F32)).contiguous()
        return [wf, c1, c2]

    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        out += ln(p + ".attn_ln") + list(fused(p + ".attn", ("query", "key", "value"), (True, False, True)))
        out += lin(p + ".attn.out") + ln(p + ".mlp_ln") + lin(p + ".mlp.0") + lin(p + ".mlp.2")
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        qkv_w, qkv_b = fused(p + ".attn", ("query", "key", "value"), (True, False, True))
        cq, fc1 = lin(p + ".cross_attn.query"), lin(p + ".mlp.0")
        out += ln(p + ".attn_ln") + [qkv_w, qkv_b]
        out += lin(p + ".attn.out") + ln(p + ".cross_attn_ln") + cq
        out += list(fused(p + ".cross_attn", ("key", "value"), (False, True)))
        out += lin(p + ".cross_attn.out") + ln(p + ".mlp_ln") + fc1 + lin(p + ".mlp.2")
        out += folded(qkv_w, qkv_b, p + ".attn_ln") + folded(cq[0], cq[1], p + ".cross_attn_ln")
        out += folded(fc1[0], fc1[1], p + ".mlp_ln")
    return out


class _Encoder:
    """Callable stand-in for `model.encoder` (reference AudioEncoder, model.py:174-204)."""

    def __init__(self, model: "Whisper"):
        self._m = model

    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        return self._m.embed_audio(mel)

    forward = __call__


class _Decoder:
    """Callable stand-in for `model.decoder` (reference TextDecoder.forward, model.py:227-249): the un-cached
    forward `decoder(tokens, audio_features) -> logits`.  The incremental path with a kv-cache lives inside the
    device session (csrc/engine.cu), so a non-empty `kv_cache` dict - the reference's hook protocol - is refused."""

    def __init__(self, model: "Whisper"):
        self._m = model

    def __call__(self, x: torch.Tensor, xa: torch.Tensor, kv_cache: Optional[dict] = None) -> torch.Tensor:
        if kv_cache:
            raise NotImplementedError("the kv-cache is resident inside the decoder session; use model.decode() / "
                                      "whisper_b200.decoding.DecoderSession for incremental decoding")
        return self._m.logits(x, xa)

    forward = __call__


class Whisper:
    def __init__(self, dims: ModelDimensions, state_dict: Optional[dict] = None, device="cuda",
                 dtype: torch.dtype = torch.float16):
        self.dims = dims
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise WhisperB200Error("whisper_b200 runs on CUDA devices only (no CPU path)")
        self.dtype = dtype
        self._handle = c_void_p(0)
        self._tensors = []
        self._workspace = None
        # decoder sessions parked between decode() calls, most recently used last (decoding.DecoderSession.close)
        self._sessions = {}
        self.timing = None                        # dict: decode() appends CUDA events per phase (bench.py)
        self.session_cache_entries = 4            # 0 disables the reuse
        self.session_cache_bytes = 64 << 30
        self.encoder = _Encoder(self)
        self.decoder = _Decoder(self)
        # default alignment heads: the last half of the decoder layers (model.py:268-276)
        heads = torch.zeros(dims.n_text_layer, dims.n_text_head, dtype=torch.bool)
        heads[dims.n_text_layer // 2:] = True
        self.alignment_heads = heads.to_sparse()
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- weights
    def load_state_dict(self, state_dict: dict):
        with torch.cuda.device(self._device):
            self._tensors = pack_weights(state_dict, self.dims, self._device, self.dtype)
            n = int(lib().wb200_model_num_tensors((ctypes.c_int32 * 10)(*self.dims.as_list())))
            if n != len(self._tensors):
                raise WhisperB200Error(f"weight packing produced {len(self._tensors)} tensors, library expects {n}")
            arr = (c_void_p * n)(*[t.data_ptr() for t in self._tensors])
            self.clear_sessions()
            if self._handle:
                lib().wb200_model_destroy(self._handle)
            h = c_void_p(0)
            check(lib().wb200_model_create((ctypes.c_int32 * 10)(*self.dims.as_list()), c_int(dtype_code(self.dtype)),
                                           arr, c_int(n), ctypes.byref(h)), "wb200_model_create")
            self._handle = h
        return self

    def set_alignment_heads(self, dump: bytes):
        """model.py:278-285."""
        import base64
        import gzip

        array = np.frombuffer(gzip.decompress(base64.b85decode(dump)), dtype=bool).copy()
        mask = torch.from_numpy(array).reshape(self.dims.n_text_layer, self.dims.n_text_head)
        self.alignment_heads = mask.to_sparse()

    def to(self, device):
        if torch.device(device) != self._device:
            raise WhisperB200Error("moving a loaded whisper_b200 model between devices is not supported; "
                                   "pass device= to load_model")
        return self

    def eval(self):
        return self

    # ---- decoder-session reuse: every decode() of the same shape otherwise re-creates the kv arenas, 96 launch plans
    # with ~1000 tensor maps and re-captures / instantiates the CUDA graph of the decode loop (transcribe() calls
    # decode() once per 30-second window, reference transcribe.py:272-508)
    def _take_session(self, key):
        sess = self._sessions.pop(key, None)
        if sess is not None:
            sess.reset_for_reuse()
        return sess

    def _park_session(self, sess) -> bool:
        if self.session_cache_entries <= 0 or sess._cache_key in self._sessions:
            return False
        self._sessions[sess._cache_key] = sess
        def total():
            return sum(s.workspace.numel() for s in self._sessions.values())
        while len(self._sessions) > self.session_cache_entries or (len(self._sessions) > 1 and total() > self.session_cache_bytes):
            oldest = next(iter(self._sessions))
            self._sessions.pop(oldest).destroy()
        return sess._cache_key in self._sessions

    def clear_sessions(self):
        for s in list(self._sessions.values()):
            s.destroy()
        self._sessions = {}

    def __del__(self):
        try:
            self.clear_sessions()
            if self._handle:
                lib().wb200_model_destroy(self._handle)
        except Exception:
            pass

    # ---- properties of the reference object (model.py:298-308)
    @property
    def device(self):
        return self._device

    @property
    def is_multilingual(self):
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self):
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)

    # ---- workspace: one growing uint8 arena shared by encoder passes
    def _arena(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, device=self._device, dtype=torch.uint8)
        return self._workspace

    # ---- forward passes
    def embed_audio(self, mel: torch.Tensor) -> torch.Tensor:
        """AudioEncoder.forward (model.py:188-204): (B, n_mels, 3000) -> (B, 1500, d)."""
        single = mel.dim() == 2
        if single:
            mel = mel[None]
        if tuple(mel.shape[1:]) != (self.dims.n_mels, 2 * self.dims.n_audio_ctx):
            raise AssertionError("incorrect audio shape")          # model.py:197
        mel = mel.to(device=self._device, dtype=torch.float32).contiguous()
        B = mel.shape[0]
        out = torch.empty((B, self.dims.n_audio_ctx, self.dims.n_audio_state), device=self._device, dtype=self.dtype)
        with torch.cuda.device(self._device):
            nbytes = int(lib().wb200_encoder_workspace_bytes(self._handle, c_int(B)))
            ws = self._arena(nbytes)
            check(lib().wb200_encoder_forward(self._handle, ptr(mel), c_int(B), ptr(out), ptr(ws), c_size_t(ws.numel()),
                                              stream_ptr()), "wb200_encoder_forward")
        return out[0] if single else out

    def logits(self, tokens: torch.Tensor, audio_features: torch.Tensor, alignment_heads=None):
        """Un-cached decoder forward over whole token rows (reference model.py:290-291 -> :227-249):
        tokens (B, n) int, audio_features (B, 1500, d) -> fp32 logits (B, n, n_vocab).
        With `alignment_heads` ((layer, head) pairs) also returns the pre-softmax cross-attention scores
        of those heads for audio 0, fp32 [n_heads, n, 1500] (what timing.py:186-197 collects with hooks)."""
        from .decoding import DecoderSession
        from .tokenizer import get_tokenizer

        tokens = torch.as_tensor(tokens)
        if tokens.dim() == 1:
            tokens = tokens[None]
        B, n = tokens.shape
        feats = audio_features.to(device=self._device, dtype=self.dtype)
        if feats.dim() == 2:
            feats = feats[None]
        feats = feats.contiguous()
        tk = get_tokenizer(self.is_multilingual, num_languages=self.num_languages)
        cfg = dict(n_audio=B, n_group=1, beam_search=0, max_candidates=1, n_init=n, sample_begin=n, sot_index=0,
                   eot=tk.eot, no_speech=-1, no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
                   suppress_blank=0, timestamp_rules=0, max_initial_timestamp_index=-1, all_logits=1)
        sess = DecoderSession(self, cfg, (), ())
        try:
            sess.set_audio(feats)
            qk = sess.set_alignment(alignment_heads) if alignment_heads is not None and len(alignment_heads) else None
            sess.prefill(tokens.cpu().numpy().astype(np.int32))
            out = sess.get_logits(B * n).reshape(B, n, self.dims.n_vocab).clone()
            if qk is not None:
                torch.cuda.current_stream(self._device).synchronize()
        finally:
            sess.close()
        return out if alignment_heads is None else (out, qk)

    def forward(self, mel: torch.Tensor, tokens: torch.Tensor):
        return self.logits(tokens, self.embed_audio(mel))

    __call__ = forward

    def install_kv_cache_hooks(self, cache=None):
        raise NotImplementedError("the kv-cache is resident inside the decoder session (csrc/engine.cu); "
                                  "there are no nn.Module hooks to install")

    # bound like the reference (model.py:343-345)
    def detect_language(self, mel, tokenizer=None):
        from .decoding import detect_language

        return detect_language(self, mel, tokenizer)

    def decode(self, mel, options=None, **kwargs):
        from .decoding import DecodingOptions, decode

        return decode(self, mel, options if options is not None else DecodingOptions(), **kwargs)

    def transcribe(self, audio, **kwargs):
        from .transcribe import transcribe

        return transcribe(self, audio, **kwargs)
