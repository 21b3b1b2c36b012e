"""CPU tests of the HF-format checkpoint loader on synthetic directories written in the formats the reference
consumes (sharded safetensors / .bin, PEFT LoRA adapter + non_lora_trainables.bin, separate tower checkpoint)."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from oracle import weights as OW
from teochat_b200 import checkpoint as CK
from teochat_b200.config import TeoConfig
from teochat_b200.weights import TeoWeights


def _write_config(d, cfg, extra=None):
    l, v = cfg.llama, cfg.vision
    c = {"hidden_size": l.hidden_size, "intermediate_size": l.intermediate_size, "num_hidden_layers": l.num_hidden_layers,
         "num_attention_heads": l.num_attention_heads, "vocab_size": l.vocab_size, "rms_norm_eps": l.rms_norm_eps,
         "max_position_embeddings": l.max_position_embeddings, "mm_projector_type": "mlp2x_gelu", "mm_vision_select_layer": -2,
         "mm_hidden_size": v.hidden_size, "tokenizer_model_max_length": 3072,
         "vision_config": {"hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size, "num_hidden_layers": v.num_hidden_layers,
                           "num_attention_heads": v.num_attention_heads, "image_size": v.image_size, "patch_size": v.patch_size}}
    c.update(extra or {})
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(c, f)


@pytest.fixture(scope="module")
def tiny_sd():
    cfg = TeoConfig.tiny()
    return cfg, OW.make_state_dict(cfg, 99, dtype=torch.float16)


def test_merged_sharded_safetensors(tmp_path, tiny_sd):
    cfg, sd = tiny_sd
    d = str(tmp_path / "teochat-merged")
    os.makedirs(d)
    _write_config(d, cfg)
    keys = sorted(sd)
    save_file({k: sd[k].contiguous() for k in keys[::2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[1::2]}, os.path.join(d, "model-00002-of-00002.safetensors"))
    c2 = CK.read_config(d)
    assert c2.llama.hidden_size == cfg.llama.hidden_size and c2.vision.num_hidden_layers == cfg.vision.num_hidden_layers
    assert c2.tokenizer_model_max_length == 3072 and c2.vit_layers_run == cfg.vit_layers_run
    got = CK.load_state_dict(d)
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    a, b = TeoWeights.from_state_dict(got, c2, "cpu"), TeoWeights.from_state_dict(sd, cfg, "cpu")
    assert all(torch.equal(a.t[k], b.t[k]) for k in b.t) and a.blocked == b.blocked
    assert type(CK.load_tokenizer(d, cfg.llama.vocab_size)).__name__ == "StubTokenizer"


def test_lora_over_base_with_separate_tower(tmp_path, tiny_sd):
    cfg, sd = tiny_sd
    base, lora, tower = str(tmp_path / "vicuna-base"), str(tmp_path / "teochat-lora"), str(tmp_path / "LanguageBind_Image")
    for d in (base, lora, tower):
        os.makedirs(d)
    _write_config(base, cfg)
    _write_config(lora, cfg)
    llm = {k: v for k, v in sd.items() if not k.startswith(CK.VIT) and "mm_projector" not in k}
    torch.save(llm, os.path.join(base, "pytorch_model.bin"))                                # .bin base without tower / projector
    tw = {"vision_model." + k[len(CK.VIT):]: v.contiguous() for k, v in sd.items() if k.startswith(CK.VIT)}
    save_file(tw, os.path.join(tower, "model.safetensors"))                                 # LanguageBind-style names
    proj = {"base_model.model." + k: v for k, v in sd.items() if "mm_projector" in k}       # builder.py:66 strips these prefixes
    torch.save(proj, os.path.join(lora, "non_lora_trainables.bin"))
    r, alpha = 4, 8
    g = torch.Generator().manual_seed(0)
    targets = ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj"]
    ad, want = {}, {k: v.clone() for k, v in sd.items()}
    for t in targets:
        out_f, in_f = sd[t + ".weight"].shape
        A, B = torch.randn(r, in_f, generator=g) * 0.1, torch.randn(out_f, r, generator=g) * 0.1
        ad[f"base_model.model.{t}.lora_A.weight"], ad[f"base_model.model.{t}.lora_B.weight"] = A, B
        want[t + ".weight"] = sd[t + ".weight"].float() + (alpha / r) * (B @ A)
    save_file(ad, os.path.join(lora, "adapter_model.safetensors"))
    with open(os.path.join(lora, "adapter_config.json"), "w") as f:
        json.dump({"r": r, "lora_alpha": alpha, "target_modules": ["q_proj", "down_proj"]}, f)
    got = CK.load_state_dict(lora, model_base=base, tower_path=tower)
    assert set(got) == set(sd)
    for k in sd:
        assert torch.allclose(got[k].float(), want[k].float(), atol=0, rtol=0) or k.replace(".weight", "") in targets and \
            torch.allclose(got[k].float(), want[k].float(), atol=1e-6), k
    with pytest.raises(KeyError):
        CK.merge_lora_({}, ad, 1.0)


def test_peft_wrapped_tower_is_merged(tmp_path, tiny_sd):
    """The reference PEFT-wraps the tower's encoder (modeling_image.py:773-792): q/k/v/out_proj appear as
    …encoder.base_model.model.layers.N.self_attn.X.base_layer.weight + lora_A/B.default.weight, and the adapter runs unmerged at
    lora_alpha / lora_r (class defaults 16 / 2).  The loader must map the names and merge W += 8·B·A — both for a tower stored
    inside the merged LLaVA checkpoint and for a separate LanguageBind_Image directory with its own config."""
    cfg, sd = tiny_sd
    g = torch.Generator().manual_seed(1)
    r = 2
    want = {k: v.clone() for k, v in sd.items()}

    def wrap(prefix_in, prefix_out, scaling):
        out = {}
        for k, v in sd.items():
            if not k.startswith(CK.VIT):
                continue
            rest = k[len(CK.VIT):]
            if rest.startswith("encoder.layers."):
                rest = "encoder.base_model.model." + rest[len("encoder."):]
                parts = rest.split(".")
                if parts[-2] in ("q_proj", "k_proj", "v_proj", "out_proj"):
                    mod = ".".join(parts[:-1])
                    out[prefix_out + mod + ".base_layer." + parts[-1]] = v.contiguous()
                    if parts[-1] == "weight":
                        A, B = torch.randn(r, v.shape[1], generator=g) * 0.05, torch.randn(v.shape[0], r, generator=g) * 0.05
                        out[prefix_out + mod + ".lora_A.default.weight"], out[prefix_out + mod + ".lora_B.default.weight"] = A, B
                        want[k] = sd[k].float() + scaling * (B @ A)
                    continue
            out[prefix_out + rest] = v.contiguous()
        return out

    # (a) tower inside the LLaVA checkpoint, no lora fields in the config -> class defaults 16/2
    d = str(tmp_path / "teochat-merged-peft-tower")
    os.makedirs(d)
    _write_config(d, cfg)
    tw = wrap(CK.VIT, CK.VIT, 8.0)
    rest = {k: v.contiguous() for k, v in sd.items() if not k.startswith(CK.VIT)}
    save_file({**rest, **tw}, os.path.join(d, "model.safetensors"))
    got = CK.load_state_dict(d)
    assert set(got) == set(sd)
    for k in sd:
        assert torch.allclose(got[k].float(), want[k].float(), atol=1e-6, rtol=0), k
    TeoWeights.from_state_dict(got, cfg, "cpu")

    # (b) separate tower directory whose config names lora_r / lora_alpha
    want = {k: v.clone() for k, v in sd.items()}
    base, tower = str(tmp_path / "llava-notower"), str(tmp_path / "LanguageBind_Image_peft")
    os.makedirs(base)
    os.makedirs(tower)
    _write_config(base, cfg)
    save_file(rest, os.path.join(base, "model.safetensors"))
    with open(os.path.join(tower, "config.json"), "w") as f:
        json.dump({"vision_config": {"lora_r": r, "lora_alpha": 4}}, f)
    save_file(wrap(CK.VIT, "vision_model.", 2.0), os.path.join(tower, "model.safetensors"))
    got = CK.load_state_dict(base, tower_path=tower)
    assert set(got) == set(sd)
    for k in sd:
        assert torch.allclose(got[k].float(), want[k].float(), atol=1e-6, rtol=0), k
    assert CK.tower_lora_scaling(tower) == 2.0 and CK.tower_lora_scaling(None, base) == 8.0


def test_rejects_unsupported(tmp_path, tiny_sd):
    cfg, _ = tiny_sd
    d = str(tmp_path / "teochat-gqa")
    os.makedirs(d)
    _write_config(d, cfg, {"num_key_value_heads": 1})
    with pytest.raises(NotImplementedError):
        CK.read_config(d)
    with pytest.raises(FileNotFoundError):
        list(CK.iter_shards(d))
