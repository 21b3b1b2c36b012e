"""Host-side token / image utilities of the hot path (mirror of videollava/mm_utils.py:10-104)."""
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


def load_image_from_base64(image):
    """mm_utils.py:10-11."""
    import base64
    from io import BytesIO

    from PIL import Image
    return Image.open(BytesIO(base64.b64decode(image)))


def expand2square(pil_img, background_color):
    """Pad a PIL image to a square on a ``background_color`` canvas, the picture centred along its short side with the odd
    pixel after it (mm_utils.py:14-25; used when ``image_aspect_ratio == 'pad'``)."""
    from PIL import Image
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def process_images(images, image_processor, model_cfg):
    """mm_utils.py:28-40: with ``model_cfg.image_aspect_ratio == 'pad'`` every image is squared on the processor's mean colour and
    preprocessed on its own (stacked when the shapes agree); otherwise the processor takes the whole list."""
    if getattr(model_cfg, "image_aspect_ratio", None) != "pad":
        return image_processor(images, return_tensors="pt")["pixel_values"]
    fill = tuple(int(c * 255) for c in image_processor.image_mean)
    out = [image_processor.preprocess(expand2square(im, fill), return_tensors="pt")["pixel_values"][0] for im in images]
    if all(x.shape == out[0].shape for x in out):
        return torch.stack(out, dim=0)
    return out


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
        # offset == 0 slices `[:, -0:]` = the WHOLE sequence, prompt included — the reference's behaviour (mm_utils.py:94), kept
        text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(k in text for k in self.keywords)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i:i + 1], scores) for i in range(output_ids.shape[0]))

    def is_eos_only(self, eos_token_id: int) -> bool:
        """True when the criterion reduces to ``last id == eos`` (the eval path's ["</s>"])."""
        return all(k.numel() == 1 and int(k[0]) == eos_token_id for k in self.keyword_ids) \
            and len(self.keyword_ids) > 0
