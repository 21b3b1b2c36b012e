"""Host-side token utilities of the hot path (mirror of videollava/mm_utils.py:43-104)."""
from __future__ import annotations

from typing import List, Sequence

import torch

from .constants import IMAGE_TOKEN_INDEX


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX,
                          return_tensors=None):
    """``<image>`` → ``image_token_index`` tokenisation (mm_utils.py:43-62).

    The prompt is split on ``<image>``; every chunk is tokenised on its own (each chunk
    therefore starts with BOS if the tokenizer prepends one); a single BOS is kept at the
    front and one ``image_token_index`` is placed between consecutive chunks.
    """
    chunks: List[List[int]] = [list(tokenizer(c).input_ids) for c in prompt.split("<image>")]
    has_bos = bool(chunks) and bool(chunks[0]) and chunks[0][0] == tokenizer.bos_token_id
    skip = 1 if has_bos else 0
    ids: List[int] = [chunks[0][0]] if has_bos else []
    for i, chunk in enumerate(chunks):
        if i > 0:
            ids.append(image_token_index)
        ids.extend(chunk[skip:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def get_model_name_from_path(model_path: str) -> str:
    """mm_utils.py:65-71."""
    parts = model_path.strip("/").split("/")
    if parts[-1].startswith("checkpoint-"):
        return parts[-2] + "_" + parts[-1]
    return parts[-1]


class KeywordsStoppingCriteria:
    """Stop when the generated tail equals a keyword's ids or its decoded text contains the
    keyword (mm_utils.py:73-104).  Same ``__call__(output_ids, scores)`` protocol as HF's
    ``StoppingCriteria``; the engine additionally recognises the common ``["</s>"]`` case and
    evaluates it on the device as ``last_id == eos``."""

    def __init__(self, keywords: Sequence[str], tokenizer, input_ids: torch.Tensor):
        self.keywords = list(keywords)
        self.keyword_ids = []
        self.max_keyword_len = 0
        for keyword in self.keywords:
            ids = list(tokenizer(keyword).input_ids)
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        for kid in self.keyword_ids:
            kid = kid.to(output_ids.device)
            if output_ids.shape[1] >= kid.shape[0] and bool((output_ids[0, -kid.shape[0]:] == kid).all()):
                return True
        if offset <= 0:
            return False
        text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(k in text for k in self.keywords)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i:i + 1], scores) for i in range(output_ids.shape[0]))

    def is_eos_only(self, eos_token_id: int) -> bool:
        """True when the criterion reduces to ``last id == eos`` (the eval path's ["</s>"])."""
        return all(k.numel() == 1 and int(k[0]) == eos_token_id for k in self.keyword_ids) \
            and len(self.keyword_ids) > 0
