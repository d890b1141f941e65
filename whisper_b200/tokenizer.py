"""Token-id side of the reference tokenizer (whisper/tokenizer.py).

The decoder kernels consume INTEGER ids only (eot, sot, timestamp_begin, the suppress lists ...);
the BPE string <-> id codec is CPU plumbing that SURVEY.md section 2 leaves out of scope.  This module
therefore reconstructs every id the hot path needs from the vocabulary size (tokenizer.py:340-355)
plus a small integer table (assets/token_ids.json: non-speech suppress ids, the id of " ", language
codes - data dumped from the reference by oracle/make_golden.py), and plugs in a real BPE codec only
if one is available at run time:

  * `tiktoken` importable AND a rank file (gpt2.tiktoken / multilingual.tiktoken) found in
    $WHISPER_B200_VOCAB_DIR or in an installed `whisper` package's assets directory.

Without it `encode()` raises and `decode()` renders ids as `<|id|>` placeholders, which is enough
for synthetic-weight benchmarking and for every parity test (they compare token ids).
"""
from __future__ import annotations

import base64
import json
import os
import string
from dataclasses import dataclass, field
from functools import cached_property, lru_cache
from typing import Dict, List, Optional, Tuple

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@lru_cache(maxsize=None)
def _id_table() -> dict:
    with open(os.path.join(_ASSETS, "token_ids.json")) as f:
        return json.load(f)


LANGUAGE_CODES: Tuple[str, ...] = tuple(_id_table()["languages"])
# language name / alias -> code (reference tokenizer.py:114-128, TO_LANGUAGE_CODE): data table dumped by oracle/make_golden.py
TO_LANGUAGE_CODE: Dict[str, str] = dict(_id_table().get("language_names", {}))


def _find_rank_file(name: str) -> Optional[str]:
    cands = []
    if os.environ.get("WHISPER_B200_VOCAB_DIR"):
        cands.append(os.path.join(os.environ["WHISPER_B200_VOCAB_DIR"], f"{name}.tiktoken"))
    try:
        import importlib.util

        spec = importlib.util.find_spec("whisper")
        if spec and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], "assets", f"{name}.tiktoken"))
    except Exception:
        pass
    for c in cands:
        if os.path.exists(c):
            return c
    return None


@lru_cache(maxsize=None)
def _bpe(name: str, num_languages: int):
    """tiktoken.Encoding with the reference's special-token layout (tokenizer.py:329-367), or None."""
    path = _find_rank_file(name)
    if path is None:
        return None
    try:
        import tiktoken
    except ImportError:
        return None
    ranks = {base64.b64decode(tok): int(rank) for tok, rank in (line.split() for line in open(path) if line)}
    n = len(ranks)
    specials = ["<|endoftext|>", "<|startoftranscript|>",
                *[f"<|{lang}|>" for lang in LANGUAGE_CODES[:num_languages]],
                "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)]]
    special_tokens = {}
    for tok in specials:
        special_tokens[tok] = n
        n += 1
    return tiktoken.Encoding(
        name=os.path.basename(path), explicit_n_vocab=n,
        pat_str=r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
        mergeable_ranks=ranks, special_tokens=special_tokens)


@dataclass
class Tokenizer:
    """Quick access to the special-token ids (reference tokenizer.py:131-327)."""

    multilingual: bool
    num_languages: int
    language: Optional[str] = None
    task: Optional[str] = None
    sot_sequence: Tuple[int, ...] = ()
    encoding: object = None          # tiktoken.Encoding or None

    def __post_init__(self):
        seq = [self.sot]
        if self.language is not None:
            seq.append(self.sot + 1 + LANGUAGE_CODES[: self.num_languages].index(self.language))
        if self.task is not None:
            seq.append(self.transcribe if self.task == "transcribe" else self.translate)
        self.sot_sequence = tuple(seq)

    # ---- ids (tokenizer.py:160-209); layout: BPE ranks, eot, sot, languages, then 6 control ids
    @cached_property
    def _base(self) -> int:
        return 50257 if self.multilingual else 50256

    @cached_property
    def eot(self) -> int:
        return self._base

    @cached_property
    def sot(self) -> int:
        return self._base + 1

    @cached_property
    def translate(self) -> int:
        return self.sot + 1 + self.num_languages

    @cached_property
    def transcribe(self) -> int:
        return self.translate + 1

    @cached_property
    def sot_lm(self) -> int:
        return self.translate + 2

    @cached_property
    def sot_prev(self) -> int:
        return self.translate + 3

    @cached_property
    def no_speech(self) -> int:
        return self.translate + 4

    @cached_property
    def no_timestamps(self) -> int:
        return self.translate + 5

    @cached_property
    def timestamp_begin(self) -> int:
        return self.translate + 6

    @cached_property
    def n_vocab(self) -> int:
        return self.timestamp_begin + 1501

    @cached_property
    def language_token(self) -> int:
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.to_language_token(self.language)

    def to_language_token(self, language: str) -> int:
        codes = LANGUAGE_CODES[: self.num_languages]
        if language in codes:
            return self.sot + 1 + codes.index(language)
        raise KeyError(f"Language {language} not found in tokenizer.")

    @cached_property
    def all_language_tokens(self) -> Tuple[int, ...]:
        return tuple(range(self.sot + 1, self.sot + 1 + self.num_languages))

    @cached_property
    def all_language_codes(self) -> Tuple[str, ...]:
        return tuple(LANGUAGE_CODES[: self.num_languages])

    @cached_property
    def sot_sequence_including_notimestamps(self) -> Tuple[int, ...]:
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        """tokenizer.py:241-276 (ids only: the symbol list is encoded once by the reference)."""
        return tuple(_id_table()["multilingual" if self.multilingual else "gpt2"]["non_speech_tokens"])

    @cached_property
    def blank_tokens(self) -> Tuple[int, ...]:
        """encode(" "), as used by SuppressBlank (decoding.py:430)."""
        return tuple(_id_table()["multilingual" if self.multilingual else "gpt2"]["blank"])

    # ---- string codec (only with a BPE rank file)
    def encode(self, text, **kwargs) -> List[int]:
        if text == " ":
            return list(self.blank_tokens)
        if self.encoding is None:
            raise RuntimeError(
                "no BPE vocabulary available: set WHISPER_B200_VOCAB_DIR to a directory holding "
                "gpt2.tiktoken / multilingual.tiktoken, or pass prompts / prefixes as token-id lists")
        return self.encoding.encode(text, **kwargs)

    def decode(self, token_ids: List[int], **kwargs) -> str:
        token_ids = [t for t in token_ids if t < self.timestamp_begin]
        if self.encoding is None:
            return "".join(f"<|{t}|>" for t in token_ids)
        return self.encoding.decode(token_ids, **kwargs)

    def decode_with_timestamps(self, token_ids: List[int], **kwargs) -> str:
        if self.encoding is None:
            return "".join(f"<|{t}|>" for t in token_ids)
        return self.encoding.decode(token_ids, **kwargs)

    def split_to_word_tokens(self, tokens: List[int]):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]):
        """tokenizer.py:287-311."""
        full = self.decode_with_timestamps(tokens)
        bad = "�"
        words, word_tokens, cur, offset = [], [], [], 0
        for token in tokens:
            cur.append(token)
            decoded = self.decode_with_timestamps(cur)
            if bad not in decoded or full[offset + decoded.index(bad)] == bad:
                words.append(decoded)
                word_tokens.append(cur)
                cur = []
                offset += len(decoded)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]):
        """tokenizer.py:313-326."""
        subwords, subword_tokens_list = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for subword, subword_tokens in zip(subwords, subword_tokens_list):
            special = subword_tokens[0] >= self.eot
            with_space = subword.startswith(" ")
            punctuation = subword.strip() in string.punctuation
            if special or with_space or punctuation or len(words) == 0:
                words.append(subword)
                word_tokens.append(subword_tokens)
            else:
                words[-1] = words[-1] + subword
                word_tokens[-1].extend(subword_tokens)
        return words, word_tokens


@lru_cache(maxsize=None)
def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language: Optional[str] = None,
                  task: Optional[str] = None) -> Tokenizer:
    """Reference tokenizer.py:370-395."""
    if language is not None:
        language = language.lower()
        if language not in LANGUAGE_CODES:
            if language in TO_LANGUAGE_CODE:                    # tokenizer.py:376-378: "english", "castilian", ...
                language = TO_LANGUAGE_CODE[language]
            else:
                raise ValueError(f"Unsupported language: {language}")
    if multilingual:
        name = "multilingual"
        language = language or "en"
        task = task or "transcribe"
    else:
        name = "gpt2"
        language = None
        task = None
    return Tokenizer(multilingual=multilingual, num_languages=num_languages, language=language, task=task,
                     encoding=_bpe(name, num_languages))
