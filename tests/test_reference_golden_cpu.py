"""Host glue and the CPU oracle against outputs of the REFERENCE's own code (tests/golden/reference_path.*, produced by
tests/golden/make_reference_golden.py from /root/reference: run_inference_single, conv_templates, tokenizer_image_token,
the torchvision transform chain, CLIPVisionTransformer, LanguageBindImageTower.feature_select, build_vision_projector,
prepare_inputs_labels_for_multimodal).  This is what pins the oracle to the reference rather than to itself."""
import numpy as np
import pytest
import torch

import refgolden
from oracle import model as OM
from oracle import weights as OW
from teochat_b200.constants import IMAGE_TOKEN_INDEX
from teochat_b200.eval.inference import build_prompt
from teochat_b200.mm_utils import tokenizer_image_token
from teochat_b200.processor import TeoImageProcessor
from teochat_b200.tokenizer import StubTokenizer


@pytest.fixture(scope="module")
def ref():
    cfg, meta, arrays = refgolden.load()
    return cfg, meta, arrays, OW.make_state_dict(cfg, meta["seed"])


def test_prompt_and_token_ids_equal_reference(ref):
    cfg, meta, arrays, _ = ref
    for ci, case in enumerate(meta["cases"]):
        prompt, paths, stop = build_prompt(case["inp"], [f"img{k}" for k in range(len(case["images"]))], "v1", case["timestamps"],
                                           case["prompt_strategy"], case["chronological_prefix"])
        assert prompt == case["prompt"]
        assert paths == [f"img{k}" for k in case["frame_order"]] and stop == "</s>"
        ids = tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size), IMAGE_TOKEN_INDEX)
        assert ids == arrays[f"input_ids_{ci}"].tolist()
        assert ids.count(IMAGE_TOKEN_INDEX) == len(case["images"]) and ids[0] == 1 and ids.count(1) == 1


def test_processor_equals_reference_transform(ref):
    """ToTensor → Resize(224, bicubic) → CenterCrop → Normalize as torchvision runs it for the reference, then the
    reference's fp16 cast (inference.py:53): the host processor reproduces every fp16 value."""
    _, meta, arrays, _ = ref
    proc = TeoImageProcessor(224)
    for ci, case in enumerate(meta["cases"]):
        got = torch.cat([proc.preprocess(im)["pixel_values"] for im in refgolden.case_images(case)])
        want = torch.from_numpy(arrays[f"pixel_values_f16_{ci}"])
        assert got.shape == want.shape
        assert torch.equal(got.to(torch.float16), want)


def test_oracle_equals_reference_tower_projector_splice_and_tokens(ref):
    cfg, meta, arrays, sd = ref
    st = meta["stride"]
    for ci, case in enumerate(meta["cases"]):
        px = torch.from_numpy(arrays[f"pixel_values_f16_{ci}"]).float()
        ids = arrays[f"input_ids_{ci}"].tolist()
        feats = OM.vit_features(sd, cfg, px)
        want = arrays[f"tower_{ci}"]
        assert np.abs(feats.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
        proj = OM.projector(sd, cfg, feats)
        want = arrays[f"projected_{ci}"]
        assert np.abs(proj.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
        emb = OM.splice(sd, cfg, ids, proj)
        assert [1, *emb.shape] == arrays[f"embeds_shape_{ci}"].tolist()
        want = arrays[f"inputs_embeds_{ci}"]
        assert np.abs(emb.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
        toks, logits = OM.generate_greedy(sd, cfg, ids, px, meta["max_new"], policy="fp32", return_logits=True)
        want = arrays[f"logits_{ci}"]
        assert toks == arrays[f"tokens_{ci}"].tolist()
        assert np.abs(logits.numpy() - want).max() <= 1e-4 * np.abs(want).max()


def test_splice_truncation_equals_reference(ref):
    """the reference's prepare_inputs_labels_for_multimodal cuts the spliced sequence at
    config.tokenizer_model_max_length (llava_arch.py:296-299) — after the image features have been spliced in."""
    import copy
    cfg, meta, arrays, sd = ref
    t = meta["truncation"]
    ci, st = t["case"], meta["stride"]
    cfg2 = copy.deepcopy(cfg)
    cfg2.tokenizer_model_max_length = t["tokenizer_model_max_length"]
    px = torch.from_numpy(arrays[f"pixel_values_f16_{ci}"]).float()
    emb = OM.splice(sd, cfg2, arrays[f"input_ids_{ci}"].tolist(), OM.encode_images(sd, cfg2, px))
    assert [1, *emb.shape] == arrays["trunc_embeds_shape"].tolist() == [1, t["tokenizer_model_max_length"], cfg.llama.hidden_size]
    want = arrays["trunc_inputs_embeds"]
    assert np.abs(emb.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
    # and it is a prefix of the untruncated splice
    full = arrays[f"inputs_embeds_{ci}"]
    assert arrays[f"embeds_shape_{ci}"][1] > t["tokenizer_model_max_length"] and full.shape[0] > want.shape[0]


def test_product_splice_plan_reproduces_reference_embeds(ref):
    """The product's integer splice plan (TeoModel.plan_splice — pure host code, called here without a GPU) gathers exactly
    the rows the reference's prepare_inputs_labels_for_multimodal concatenates: applying the plan to the oracle's embedding table
    and projector output gives the reference's inputs_embeds, for every fixture case, ragged in one batch, and truncated."""
    import copy
    from types import SimpleNamespace

    from teochat_b200.engine import TeoModel
    cfg, meta, arrays, sd = ref
    st = meta["stride"]
    E = sd["model.embed_tokens.weight"].float()
    ids_all = [arrays[f"input_ids_{ci}"].tolist() for ci in range(len(meta["cases"]))]
    n_img = [len(c["images"]) for c in meta["cases"]]
    feats = torch.cat([OM.encode_images(sd, cfg, torch.from_numpy(arrays[f"pixel_values_f16_{ci}"]).float()) for ci in range(len(ids_all))])
    flat = feats.reshape(-1, cfg.llama.hidden_size)                 # [all frames * 256, h]: the plan's image rows index this
    srcs, lens = TeoModel.plan_splice(SimpleNamespace(cfg=cfg), ids_all, n_img)
    for ci, (src, n) in enumerate(zip(srcs, lens)):
        assert [1, n, cfg.llama.hidden_size] == arrays[f"embeds_shape_{ci}"].tolist()
        src = torch.from_numpy(src)
        emb = torch.where((src >= 0)[:, None], E[src.clamp_min(0)], flat[(-(src + 1)).clamp_min(0)])
        want = arrays[f"inputs_embeds_{ci}"]
        assert np.abs(emb.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
    t = meta["truncation"]
    cfg2 = copy.deepcopy(cfg)
    cfg2.tokenizer_model_max_length = t["tokenizer_model_max_length"]
    srcs, lens = TeoModel.plan_splice(SimpleNamespace(cfg=cfg2), [ids_all[t["case"]]], [n_img[t["case"]]])
    assert lens == [t["tokenizer_model_max_length"]]
    src = torch.from_numpy(srcs[0])
    one = OM.encode_images(sd, cfg, torch.from_numpy(arrays[f"pixel_values_f16_{t['case']}"]).float()).reshape(-1, cfg.llama.hidden_size)
    emb = torch.where((src >= 0)[:, None], E[src.clamp_min(0)], one[(-(src + 1)).clamp_min(0)])
    want = arrays["trunc_inputs_embeds"]
    assert np.abs(emb.flatten()[::st].numpy() - want).max() <= 2e-5 * np.abs(want).max()
