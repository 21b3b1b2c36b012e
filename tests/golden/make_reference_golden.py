"""Generates tests/golden/reference_path.npz + reference_path.json by running the REFERENCE's OWN code for the hot
path (imported from /root/reference in this container only; nothing here can run on the GPU box):

    videollava/eval/inference.py           run_inference_single, replace_video_token            (unmodified, called as is)
    videollava/conversation.py             conv_templates["v1"]
    videollava/mm_utils.py                 tokenizer_image_token, KeywordsStoppingCriteria
    …/languagebind/image/processing_image.py   LanguageBindImageProcessor / get_image_transform (torchvision)
    …/languagebind/image/modeling_image.py     CLIPVisionTransformer (the tower's arithmetic)
    …/languagebind/__init__.py             LanguageBindImageTower.forward / feature_select
    videollava/model/multimodal_projector/builder.py   build_vision_projector("mlp2x_gelu")
    videollava/model/llava_arch.py         LlavaMetaForCausalLM.encode_images / prepare_inputs_labels_for_multimodal

What is NOT the reference here, and why:
  * third-party packages the reference imports at module scope but that are absent offline and unused on this path
    (peft, decord, pytorchvideo, torchaudio, cv2) are mocked; `_expand_mask` (removed from transformers 5.5, used only by
    the CLIP text side) is a dummy.  `videollava/__init__.py` and `videollava/model/__init__.py` are bypassed (they
    import llava_llama.py, which needs transformers==4.31 internals) — sub-modules are imported by path.
  * the language model: llava_llama.py cannot load on transformers 5.5, so `model.generate` is a 30-line
    driver below = the reference's prepare_inputs_labels_for_multimodal → installed HF `LlamaForCausalLM(eager)` greedy
    loop with the reference's own stopping criteria object.  LLaMA arithmetic is HF's in either case (third party).
  * the tokenizer: the LLaMA sentencepiece model is not available offline → teochat_b200.tokenizer.StubTokenizer
    (duck type only; the -200 splicing under test is the reference's tokenizer_image_token).
  * the reference casts frames to fp16 (inference.py:53) for its fp16 model; the driver upcasts them back to fp32 for the
    fp32 modules, so the recorded pixel values are the fp16-ROUNDED processor outputs.
Weights: oracle/weights.make_state_dict (hash init) on a small config with the real 224² / patch-14 geometry
(257 tokens per frame) so the reference processor (hard-coded 224) applies.

Run: python tests/golden/make_reference_golden.py
"""
import importlib
import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SEED = 777
MAX_NEW = 8
STRIDE = 7                     # feature tensors are stored every STRIDE-th element (flattened) to keep the fixture small


def ref_config():
    """Small widths, REAL image geometry (224², patch 14 → 256 patches + CLS)."""
    sys.path.insert(0, ROOT)
    from teochat_b200.config import LlamaConfig, TeoConfig, VisionConfig
    v = VisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=224, patch_size=14)
    l = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                    max_position_embeddings=2048)
    return TeoConfig(vision=v, llama=l, kv_page_size=16)


def import_reference():
    def stub_pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    for name in list(sys.modules):
        if name == "videollava" or name.startswith("videollava."):
            del sys.modules[name]                     # never the repo's own shim package
    stub_pkg("videollava", REF + "/videollava")
    stub_pkg("videollava.model", REF + "/videollava/model")
    for name in ["decord", "pytorchvideo", "pytorchvideo.data", "pytorchvideo.data.encoded_video", "pytorchvideo.transforms",
                 "torchaudio", "torchaudio.compliance", "torchaudio.compliance.kaldi", "cv2", "peft"]:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
    import transformers.models.clip.modeling_clip as mc
    if not hasattr(mc, "_expand_mask"):
        mc._expand_mask = lambda *a, **k: None
    mods = SimpleNamespace()
    mods.lb = importlib.import_module("videollava.model.multimodal_encoder.languagebind")
    mods.mi = importlib.import_module("videollava.model.multimodal_encoder.languagebind.image.modeling_image")
    mods.ci = importlib.import_module("videollava.model.multimodal_encoder.languagebind.image.configuration_image")
    mods.pi = importlib.import_module("videollava.model.multimodal_encoder.languagebind.image.processing_image")
    mods.pb = importlib.import_module("videollava.model.multimodal_projector.builder")
    mods.arch = importlib.import_module("videollava.model.llava_arch")
    mods.inf = importlib.import_module("videollava.eval.inference")
    mods.mm = importlib.import_module("videollava.mm_utils")
    mods.conv = importlib.import_module("videollava.conversation")
    for m in vars(mods).values():
        assert m.__file__.startswith(REF), m.__file__
    return mods


class RecordingTokenizer:
    """StubTokenizer that remembers the text chunks it was asked to encode (→ the reference's final prompt)."""

    def __init__(self, inner):
        self.inner, self.chunks = inner, []
        self.bos_token_id, self.eos_token_id = inner.bos_token_id, inner.eos_token_id

    def __call__(self, text, **kw):
        self.chunks.append(text)
        return self.inner(text, **kw)

    def decode(self, *a, **k):
        return self.inner.decode(*a, **k)

    def batch_decode(self, *a, **k):
        return self.inner.batch_decode(*a, **k)


def build_reference_model(mods, cfg, sd):
    from transformers import LlamaConfig, LlamaForCausalLM
    v, l = cfg.vision, cfg.llama
    vc = mods.ci.CLIPVisionConfig(hidden_size=v.hidden_size, intermediate_size=v.intermediate_size, num_hidden_layers=v.num_hidden_layers,
                                  num_attention_heads=v.num_attention_heads, image_size=v.image_size, patch_size=v.patch_size,
                                  hidden_act=v.hidden_act, layer_norm_eps=v.layer_norm_eps, lora_r=0)
    vc._attn_implementation = "eager"
    vis = mods.mi.CLIPVisionTransformer(vc).eval()
    pre = "model.image_tower.image_tower."
    res = vis.load_state_dict({k[len(pre):]: t for k, t in sd.items() if k.startswith(pre)}, strict=False)
    assert not res.unexpected_keys and all("post_layernorm" in k or "position_ids" in k for k in res.missing_keys), res
    tower = mods.lb.LanguageBindImageTower.__new__(mods.lb.LanguageBindImageTower)      # skip from_pretrained (network)
    nn.Module.__init__(tower)
    tower.is_loaded, tower.select_layer, tower.select_feature = True, cfg.mm_vision_select_layer, cfg.mm_vision_select_feature
    tower.image_tower = vis
    proj = mods.pb.build_vision_projector(SimpleNamespace(mm_projector_type=cfg.mm_projector_type, mm_hidden_size=v.hidden_size,
                                                          hidden_size=l.hidden_size)).eval()
    proj.load_state_dict({k[len("model.mm_projector."):]: t for k, t in sd.items() if k.startswith("model.mm_projector.")})
    lc = LlamaConfig(hidden_size=l.hidden_size, intermediate_size=l.intermediate_size, num_hidden_layers=l.num_hidden_layers,
                     num_attention_heads=l.num_attention_heads, num_key_value_heads=l.num_attention_heads, vocab_size=l.vocab_size,
                     rms_norm_eps=l.rms_norm_eps, rope_theta=l.rope_theta, max_position_embeddings=l.max_position_embeddings,
                     attn_implementation="eager", tie_word_embeddings=False)
    lm = LlamaForCausalLM(lc).eval()
    lsd = {k: t for k, t in sd.items() if k.startswith("model.layers") or k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")}
    assert not lm.load_state_dict(lsd, strict=False).unexpected_keys

    inner = SimpleNamespace(embed_tokens=lm.model.embed_tokens, mm_projector=proj, get_image_tower=lambda: tower,
                            get_video_tower=lambda: None)

    class ReferenceGlue(mods.arch.LlavaMetaForCausalLM):
        """The reference's multimodal glue (encode_images, prepare_inputs_labels_for_multimodal — inherited, unmodified)
        over an HF LLaMA; `generate` is the stand-in for LlavaLlamaForCausalLM.generate described in the module docstring."""
        device = torch.device("cpu")
        config = SimpleNamespace(tokenizer_model_max_length=cfg.tokenizer_model_max_length, tokenizer_padding_side="right")

        def __init__(self):
            self.trace = {}

        def get_model(self):
            return inner

        def generate(self, input_ids, images, do_sample, temperature, max_new_tokens, use_cache, stopping_criteria):
            images = [im.float() for im in images]            # fp16-rounded values, fp32 arithmetic
            self.trace["pixel_values"] = torch.stack(images)
            with torch.no_grad():
                self.trace["tower"] = tower(torch.stack(images))
                self.trace["projected"] = self.encode_images(torch.stack(images))
                _, pos, _, _, emb, _ = self.prepare_inputs_labels_for_multimodal(input_ids, None, None, None, None, images)
                assert pos is None                            # HF then uses arange(S0)
                self.trace["inputs_embeds"] = emb
                out = lm(inputs_embeds=emb, use_cache=True)
                ids, logits = input_ids, []
                for _ in range(max_new_tokens):
                    lg = out.logits[0, -1].float()
                    logits.append(lg)
                    tok = int(lg.argmax())                    # greedy (the reference's do_sample=True is not reproducible)
                    ids = torch.cat([ids, torch.tensor([[tok]])], dim=1)
                    if any(sc(ids, None) for sc in stopping_criteria):
                        break
                    out = lm(input_ids=torch.tensor([[tok]]), past_key_values=out.past_key_values, use_cache=True)
            self.trace["logits"] = torch.stack(logits)
            return ids

    return ReferenceGlue()


CASES = [
    # inp, [(h, w, image seed)], timestamps, prompt_strategy, chronological_prefix
    ("This is a sequence of images captured at times: <video> What objects or changes can you see across the images?",
     [(300, 400, 1), (224, 224, 2)], ["2021-06-01", "2019-01-31"], "interleave", True),
    ("Classify the image <video> into one of: airport, farm, port.", [(640, 513, 3)], [], "interleave", True),
    ("times: <video> Describe.", [(224, 320, 4), (100, 150, 5), (224, 224, 6)], [], None, False),
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    cfg = ref_config()
    mods = import_reference()
    from oracle import weights as OW
    from teochat_b200.tokenizer import StubTokenizer
    sd = OW.make_state_dict(cfg, SEED)
    model = build_reference_model(mods, cfg, sd)
    try:
        processor = mods.pi.LanguageBindImageProcessor(SimpleNamespace(vision_config=SimpleNamespace(image_size=224)))
        proc_kind = "LanguageBindImageProcessor"
    except Exception as e:                                    # ProcessorMixin.__init__ differs across transformers versions
        transform = mods.pi.get_image_transform(SimpleNamespace(vision_config=None))
        processor = SimpleNamespace(preprocess=lambda im, return_tensors: {"pixel_values": torch.stack([mods.pi.load_and_transform_image(im, transform)])})
        proc_kind = f"get_image_transform + load_and_transform_image (ProcessorMixin init failed: {type(e).__name__})"
    from PIL import Image
    meta = {"seed": SEED, "max_new": MAX_NEW, "stride": STRIDE, "processor": proc_kind, "cases": [],
            "config": {"vision": vars(cfg.vision), "llama": vars(cfg.llama), "kv_page_size": cfg.kv_page_size}}
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        for ci, (inp, images, stamps, strategy, chrono) in enumerate(CASES):
            paths = []
            for k, (h, w, s) in enumerate(images):
                arr = np.random.RandomState(s).randint(0, 256, (h, w, 3), dtype=np.uint8)
                p = os.path.join(tmp, f"c{ci}_{k}.png")
                Image.fromarray(arr).save(p)
                paths.append(p)
            tok = RecordingTokenizer(StubTokenizer(cfg.llama.vocab_size))
            out = mods.inf.run_inference_single(model, processor, tok, inp, paths, conv_mode="v1", timestamps=list(stamps),
                                                prompt_strategy=strategy, chronological_prefix=chrono, temperature=0.2,
                                                max_new_tokens=MAX_NEW)
            n_chunks = len(images) + 1
            prompt = "<image>".join(tok.chunks[:n_chunks])     # tokenizer_image_token encodes the chunks first
            ids = mods.mm.tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size), -200)
            t = model.trace
            new = t["logits"].shape[0]
            top2 = t["logits"].topk(2, -1).values
            order = list(range(len(images)))
            if stamps:
                from datetime import datetime
                order = sorted(order, key=lambda i: datetime.strptime(stamps[i], "%Y-%m-%d"))
            meta["cases"].append({"inp": inp, "images": images, "timestamps": list(stamps), "prompt_strategy": strategy,
                                  "chronological_prefix": chrono, "prompt": prompt, "output": out, "frame_order": order})
            arrays[f"input_ids_{ci}"] = np.asarray(ids, dtype=np.int64)
            arrays[f"pixel_values_f16_{ci}"] = t["pixel_values"].to(torch.float16).numpy()
            arrays[f"tower_{ci}"] = t["tower"].flatten()[::STRIDE].numpy()
            arrays[f"projected_{ci}"] = t["projected"].flatten()[::STRIDE].numpy()
            arrays[f"inputs_embeds_{ci}"] = t["inputs_embeds"].flatten()[::STRIDE].numpy()
            arrays[f"embeds_shape_{ci}"] = np.asarray(t["inputs_embeds"].shape, dtype=np.int64)
            arrays[f"logits_{ci}"] = t["logits"].numpy()
            arrays[f"tokens_{ci}"] = t["logits"].argmax(-1).numpy().astype(np.int64)
            arrays[f"margin_{ci}"] = (top2[:, 0] - top2[:, 1]).numpy()
            print(f"case {ci}: prompt {len(prompt)} chars, ids {len(ids)}, S0 {t['inputs_embeds'].shape[1]}, {new} new tokens "
                  f"{arrays[f'tokens_{ci}'].tolist()} → {out!r}; min top-2 margin {arrays[f'margin_{ci}'].min():.3g}", flush=True)
        # quirk 4 (llava_arch.py:296-299): the spliced sequence is cut at config.tokenizer_model_max_length — the reference's
        # own prepare_inputs_labels_for_multimodal on case 1 (306 positions) with the attribute set to 200
        model.config.tokenizer_model_max_length = 200
        px = torch.from_numpy(arrays["pixel_values_f16_1"]).float()
        with torch.no_grad():
            _, _, _, _, emb, _ = model.prepare_inputs_labels_for_multimodal(torch.from_numpy(arrays["input_ids_1"])[None], None, None, None,
                                                                           None, [im for im in px])
        model.config.tokenizer_model_max_length = None
        meta["truncation"] = {"case": 1, "tokenizer_model_max_length": 200}
        arrays["trunc_embeds_shape"] = np.asarray(emb.shape, dtype=np.int64)
        arrays["trunc_inputs_embeds"] = emb.flatten()[::STRIDE].numpy()
        print("truncation: inputs_embeds", tuple(emb.shape), flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_path.npz"), **arrays)
    with open(os.path.join(HERE, "reference_path.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_path.npz"), os.path.getsize(os.path.join(HERE, "reference_path.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
