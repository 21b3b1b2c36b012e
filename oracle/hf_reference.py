"""ORACLE / CPU BASELINE — test infrastructure only (imported by bench.py's CPU legs and tests/, never by the product).

The "reference HF/PyTorch CPU path" of BASELINE.md §2: the reference package itself cannot be imported here (it pins
transformers==4.31 / peft / decord), so its hot path is driven the way the reference drives it —
  run_inference_single (videollava/eval/inference.py:23-77): preprocess → model.generate(greedy)
  encode_images (llava_arch.py:137-140): tower.hidden_states[-2][:, 1:] (languagebind/__init__.py:121-129) → mlp2x_gelu
  prepare_inputs_labels_for_multimodal (llava_arch.py:251-299): the splice (oracle.model.splice, pinned to the reference's own
  output by tests/test_reference_golden_cpu.py)
  LlavaLlamaForCausalLM.forward → HF LlamaForCausalLM (llava_llama.py:88-99), all-position logits like the reference
— through the INSTALLED transformers modules (CLIPVisionModel, LlamaForCausalLM, eager attention), fp32, random-init weights
from oracle.weights (the same numbers the GPU build uses).  `kind` of this baseline is therefore "port", not "reference".
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence

import torch

from . import model as OM


def build_modules(cfg, sd: Dict[str, torch.Tensor]):
    """(clip, projector, llama) HF modules sharing storage with the fp32 state dict `sd` (no second 27 GB copy)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM
    v, l = cfg.vision, cfg.llama
    hc = CLIPVisionConfig(hidden_size=v.hidden_size, intermediate_size=v.intermediate_size, num_hidden_layers=v.num_hidden_layers,
                          num_attention_heads=v.num_attention_heads, image_size=v.image_size, patch_size=v.patch_size,
                          hidden_act=v.hidden_act, layer_norm_eps=v.layer_norm_eps, attn_implementation="eager")
    clip = CLIPVisionModel(hc).eval()
    hsd = {"vision_model." + k[len(OM.VIT):]: t.float() for k, t in sd.items() if k.startswith(OM.VIT)}
    res = clip.load_state_dict(hsd, strict=False)
    assert not res.unexpected_keys and all("post_layernorm" in k for k in res.missing_keys), res
    proj = torch.nn.Sequential(torch.nn.Linear(v.hidden_size, l.hidden_size), torch.nn.GELU(), torch.nn.Linear(l.hidden_size, l.hidden_size)).eval()
    proj.load_state_dict({k[len("model.mm_projector."):]: t.float() for k, t in sd.items() if k.startswith("model.mm_projector.")})
    lc = LlamaConfig(hidden_size=l.hidden_size, intermediate_size=l.intermediate_size, num_hidden_layers=l.num_hidden_layers,
                     num_attention_heads=l.num_attention_heads, num_key_value_heads=l.num_attention_heads, vocab_size=l.vocab_size,
                     rms_norm_eps=l.rms_norm_eps, rope_theta=l.rope_theta, max_position_embeddings=max(l.max_position_embeddings, 8192),
                     attn_implementation="eager", tie_word_embeddings=False)
    with torch.device("meta"):
        lm = LlamaForCausalLM(lc)
    lsd = {k: t.float() for k, t in sd.items() if k.startswith("model.layers") or k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")}
    res = lm.load_state_dict(lsd, strict=False, assign=True)
    assert not res.unexpected_keys and not res.missing_keys, res
    lm.model.rotary_emb = type(lm.model.rotary_emb)(config=lc)          # non-persistent inv_freq buffer: rebuild off the meta device
    return clip, proj, lm.eval()


@torch.no_grad()
def run_inference_greedy(modules, cfg, sd, input_ids: Sequence[int], frames_u8: torch.Tensor, max_new_tokens: int):
    """One example end to end; returns (new token ids, phase seconds)."""
    clip, proj, lm = modules
    t0 = time.perf_counter()
    px = OM.normalize_u8_nhwc(frames_u8)                                   # processing_image.py:18,22 (224² frames: resize/crop are identity)
    hs = clip(pixel_values=px, output_hidden_states=True).hidden_states[cfg.mm_vision_select_layer]
    feats = proj(hs[:, 1:])
    emb = OM.splice(sd, cfg, input_ids, feats)
    t1 = time.perf_counter()
    o = lm(inputs_embeds=emb[None], use_cache=True)                        # all-position logits, like the reference
    toks: List[int] = [int(o.logits[0, -1].argmax())]
    t2 = time.perf_counter()
    while len(toks) < max_new_tokens:
        o = lm(input_ids=torch.tensor([[toks[-1]]]), past_key_values=o.past_key_values, use_cache=True)
        toks.append(int(o.logits[0, -1].argmax()))
    t3 = time.perf_counter()
    return toks, {"vision_s": t1 - t0, "prefill_s": t2 - t1, "decode_s": t3 - t2, "context": int(emb.shape[0])}
