"""Static shapes of the hot path.

One dataclass mirrors the ``config.json`` attributes the reference reads with ``getattr``
(llava_arch.py:32-37,296,312; builder.py:139-149,166; configuration_image.py:181-232).
"""
from __future__ import annotations

import dataclasses
from typing import Optional


@dataclasses.dataclass
class VisionConfig:
    """CLIP ViT-L/14 @224 (configuration_image.py:181-232)."""
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    image_size: int = 224
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-5
    hidden_act: str = "quick_gelu"       # configuration_image.py:191 (config field)
    initializer_range: float = 0.02
    initializer_factor: float = 1.0

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def num_positions(self) -> int:
        return self.num_patches + 1

    @property
    def patch_dim(self) -> int:
        return self.num_channels * self.patch_size * self.patch_size


@dataclasses.dataclass
class LlamaConfig:
    """LLaMA-2-7B / Vicuna-7B-v1.5."""
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    vocab_size: int = 32000
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    initializer_range: float = 0.02
    bos_token_id: int = 1
    eos_token_id: int = 2

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclasses.dataclass
class TeoConfig:
    vision: VisionConfig = dataclasses.field(default_factory=VisionConfig)
    llama: LlamaConfig = dataclasses.field(default_factory=LlamaConfig)
    mm_projector_type: str = "mlp2x_gelu"          # multimodal_projector/builder.py:41-48
    mm_vision_select_layer: int = -2               # languagebind/__init__.py:121
    mm_vision_select_feature: str = "patch"        # languagebind/__init__.py:123
    mm_use_im_start_end: bool = False
    mm_use_im_patch_token: bool = False
    tokenizer_model_max_length: Optional[int] = None   # llava_arch.py:296 (None = no truncation)
    tokenizer_padding_side: str = "right"          # llava_arch.py:312
    kv_page_size: int = 64                         # tokens per KV page (new; no reference analogue)

    @property
    def mm_hidden_size(self) -> int:
        return self.vision.hidden_size

    @property
    def vit_layers_run(self) -> int:
        """hidden_states[select_layer] needs only this many encoder layers."""
        n = self.vision.num_hidden_layers
        sl = self.mm_vision_select_layer
        idx = sl if sl >= 0 else n + 1 + sl      # index into the (n+1)-long hidden_states tuple
        if not 0 <= idx <= n:
            raise ValueError(f"mm_vision_select_layer {sl} out of range for {n} layers")
        return idx

    @property
    def tokens_per_image(self) -> int:
        if self.mm_vision_select_feature == "patch":
            return self.vision.num_patches
        if self.mm_vision_select_feature == "cls_patch":
            return self.vision.num_positions
        raise ValueError(f"Unexpected select feature: {self.mm_vision_select_feature}")

    @staticmethod
    def full() -> "TeoConfig":
        return TeoConfig()

    @staticmethod
    def tiny() -> "TeoConfig":
        """A CPU-second-scale config with the same structure (tests / smoke)."""
        v = VisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                         num_attention_heads=2, image_size=56, patch_size=14)
        l = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                        num_attention_heads=2, vocab_size=512, max_position_embeddings=512)
        return TeoConfig(vision=v, llama=l, kv_page_size=16)
