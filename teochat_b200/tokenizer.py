"""Offline stand-in for the LLaMA sentencepiece tokenizer.

The reference loads ``AutoTokenizer.from_pretrained(model_path, use_fast=False)``
(builder.py:111); neither ``tokenizer.model`` nor network is available offline, so tests,
``smoke()`` and ``bench.py`` use this deterministic word-hash tokenizer with the same duck type
the hot path touches: ``__call__(str).input_ids`` (BOS first), ``bos_token_id``,
``eos_token_id``, ``decode``, ``batch_decode`` (mm_utils.py:44,51,94; inference.py:75).
"""
from __future__ import annotations

import re
import zlib
from types import SimpleNamespace
from typing import Iterable, List


class StubTokenizer:
    bos_token_id = 1
    eos_token_id = 2
    unk_token_id = 0
    pad_token_id = 0

    def __init__(self, vocab_size: int = 32000):
        self.vocab_size = int(vocab_size)
        self._words = {}           # id -> first word seen (so decode round-trips seen text)

    def _word_id(self, w: str) -> int:
        wid = 3 + zlib.crc32(w.encode("utf-8")) % (self.vocab_size - 3)
        self._words.setdefault(wid, w)
        return wid

    def tokenize(self, text: str) -> List[str]:
        return re.findall(r"</s>|\w+|[^\w\s]", text)

    def __call__(self, text: str, **kwargs):
        ids = [self.bos_token_id]
        for w in self.tokenize(text):
            ids.append(self.eos_token_id if w == "</s>" else self._word_id(w))
        return SimpleNamespace(input_ids=ids)

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = False) -> str:
        out = []
        for i in (int(x) for x in ids):
            if i == self.bos_token_id:
                if not skip_special_tokens:
                    out.append("<s>")
            elif i == self.eos_token_id:
                if not skip_special_tokens:
                    out.append("</s>")
            elif i < 0:
                continue                      # IMAGE_TOKEN_INDEX never reaches decode slices
            else:
                out.append(self._words.get(i, f"<{i}>"))
        return " ".join(out)

    def batch_decode(self, batch, skip_special_tokens: bool = False) -> List[str]:
        return [self.decode(row, skip_special_tokens=skip_special_tokens) for row in batch]
