"""GPU parity tests, model level: the whole B200 path (through the C-ABI, via teochat_b200.engine)
against the CPU oracle on the same seeded inputs, and against the committed golden vectors.

Floating-point bar (BASELINE.json north_star): logits within 1e-2 relative (of the row's max |logit|)
of the fp32 oracle; greedy token ids equal to the bf16-policy oracle, where a mismatch is accepted
only at a step whose oracle top-2 margin is below NEAR_TIE (two logits closer than the bf16
pipeline can resolve) — the comparison stops there because the continuations legitimately differ.
"""
import os

import numpy as np
import pytest
import torch

from teochat_b200.config import TeoConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGIT_RTOL = 1e-2
NEAR_TIE = 2e-2          # relative to the row's max |logit|


def _model(cfg, seed=1234):
    from teochat_b200.engine import TeoModel
    from teochat_b200.weights import TeoWeights
    return TeoModel(cfg, TeoWeights.from_synthetic(cfg, seed, DEV), DEV)


@pytest.fixture(scope="module")
def tiny():
    cfg = TeoConfig.tiny()
    return cfg, _model(cfg)


@pytest.fixture(scope="module")
def tiny_oracle():
    from oracle import weights as OW
    cfg = TeoConfig.tiny()
    return cfg, OW.make_state_dict(cfg, 1234)


def compare_tokens(got, want, margins, absmax, what=""):
    """ids must agree up to the first oracle near-tie; returns the number of verified steps."""
    n = min(len(got), len(want))
    for i in range(n):
        if got[i] != want[i]:
            assert margins[i] <= NEAR_TIE * absmax[i], f"{what}: token {i} differs ({got[i]} vs {want[i]}) at margin {margins[i]:.4g} / {absmax[i]:.4g}"
            return i
    assert len(got) == len(want) or n == len(want), f"{what}: length {len(got)} vs {len(want)}"
    return n


def test_synthetic_weights_match_oracle_bitexact(tiny, tiny_oracle):
    cfg, model = tiny
    _, sd = tiny_oracle
    from teochat_b200.weights import TeoWeights
    w2 = TeoWeights.from_state_dict(sd, cfg, DEV)
    for k, t in model.w.t.items():
        assert torch.equal(t, w2.t[k]), k


def test_vit_projector_vs_oracle(tiny, tiny_oracle):
    from oracle import model as OM
    from oracle import weights as OW
    cfg, model = tiny
    _, sd = tiny_oracle
    frames = OW.synthetic_frames_u8(5, cfg.vision.image_size, 3)
    px = OM.normalize_u8_nhwc(frames)
    got = model.encode_images(frames_u8=frames.to(DEV)).float().cpu()
    got_px = model.encode_images(pixel_values=px.to(DEV)).float().cpu()
    assert torch.equal(got, got_px)                      # u8 path ≡ reference float pixel_values path
    ref32 = OM.encode_images(sd, cfg, px, "fp32")
    ref16 = OM.encode_images(sd, cfg, px, "bf16")
    s = ref32.abs().max().item()
    assert (got - ref32).abs().max().item() <= 2e-2 * s     # bf16 activations through 2 ViT layers + projector
    assert (got - ref16).abs().max().item() <= 1e-2 * s     # same rounding points: only accumulation order differs


def test_generate_tiny_vs_golden(tiny):
    from oracle import weights as OW
    cfg, model = tiny
    z = np.load(os.path.join(GOLDEN, "tiny_generate.npz"))
    assert int(z["seed"]) == 1234
    ids, frames = [], []
    for i in range(int(z["n_samples"])):
        ids.append(z[f"ids_{i}"].tolist())
        nf, fs = z[f"frames_{i}"].tolist()
        frames.append(OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs))
    max_new = int(z["max_new"])
    outs, logits = model.generate_batch(ids, frames_u8=frames, max_new_tokens=max_new, return_logits=True)
    stride = int(z["logit_stride"])
    verified = []
    for i, out in enumerate(outs):
        l0 = logits[i, 0, ::stride].float().cpu().numpy()
        ref = z[f"logits0_fp32_{i}"]
        assert np.abs(l0 - ref).max() <= LOGIT_RTOL * np.abs(ref).max(), f"sample {i} step-0 logits"
        verified.append(compare_tokens(out, z[f"tokens_bf16_{i}"].tolist(), z[f"margin_bf16_{i}"], z[f"absmax_bf16_{i}"], f"sample {i}"))
    print("verified greedy steps per sample:", verified, "of", max_new)
    assert min(verified) >= 1
    # the batched ragged run must reproduce single-sample runs (independent units)
    for i in range(len(ids)):
        single = model.generate_batch([ids[i]], frames_u8=[frames[i]], max_new_tokens=max_new)[0]
        z_m, z_a = z[f"margin_bf16_{i}"], z[f"absmax_bf16_{i}"]
        compare_tokens(single, outs[i], z_m, z_a, f"single-vs-batch {i}")


def test_generate_tiny_vs_live_oracle(tiny, tiny_oracle):
    """Fresh seeds (not in the fixture) against the oracle run on the box's CPU."""
    from oracle import model as OM
    from oracle import weights as OW
    cfg, model = tiny
    _, sd = tiny_oracle
    ids = [1, 17, 99, -200, 5, 6, -200, 300, 301, 302]
    frames = OW.synthetic_frames_u8(2, cfg.vision.image_size, 99)
    px = OM.normalize_u8_nhwc(frames)
    want, wl = OM.generate_greedy(sd, cfg, ids, px, 12, policy="bf16", return_logits=True)
    _, wl32 = OM.generate_greedy(sd, cfg, ids, px, 1, policy="fp32", return_logits=True)
    got, gl = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=12, return_logits=True)
    g0 = gl[0, 0].float().cpu()
    assert (g0 - wl32[0]).abs().max().item() <= LOGIT_RTOL * wl32[0].abs().max().item()
    top2 = wl.topk(2, -1).values
    n = compare_tokens(got[0], want, (top2[:, 0] - top2[:, 1]).numpy(), wl.abs().amax(-1).numpy(), "live oracle")
    # while tokens agree, every step's logits must stay within tolerance of the bf16-policy oracle
    for s in range(n):
        assert (gl[0, s].float().cpu() - wl[s]).abs().max().item() <= LOGIT_RTOL * wl[s].abs().max().item(), s


def test_generate_odd_widths_row_major_weights_vs_live_oracle():
    """A config whose LLaMA MLP width (144) is a multiple of neither 32 nor 64: gate/up stay [gate; up] (SwiGLU as a kernel,
    prefill and decode), the LLaMA weights stay row-major (2-D TMA maps) — the layouts a checkpoint with unusual widths gets."""
    from oracle import model as OM
    from oracle import weights as OW
    cfg = TeoConfig.tiny()
    cfg.llama.intermediate_size = 144
    model = _model(cfg, 4242)
    assert not model.w.gate_up_interleaved and not model.w.blocked["llama"]
    sd = OW.make_state_dict(cfg, 4242)
    ids = [1, 17, 99, -200, 5, 6, -200, 300, 301, 302]
    frames = OW.synthetic_frames_u8(2, cfg.vision.image_size, 98)
    px = OM.normalize_u8_nhwc(frames)
    want, wl = OM.generate_greedy(sd, cfg, ids, px, 10, policy="bf16", return_logits=True)
    got, gl = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=10, return_logits=True)
    top2 = wl.topk(2, -1).values
    n = compare_tokens(got[0], want, (top2[:, 0] - top2[:, 1]).numpy(), wl.abs().amax(-1).numpy(), "odd widths")
    assert n >= 1
    for s in range(n):
        assert (gl[0, s].float().cpu() - wl[s]).abs().max().item() <= LOGIT_RTOL * wl[s].abs().max().item(), s
    assert model.generate_batch([ids], frames_u8=[frames], max_new_tokens=10) == got      # graph replay, same ids
    del model
    torch.cuda.empty_cache()


def _full_width(depth_llama, depth_vit):
    cfg = TeoConfig.full()
    cfg.llama.num_hidden_layers = depth_llama
    cfg.vision.num_hidden_layers = depth_vit + 1        # select_layer -2 → depth_vit layers executed
    return cfg


@pytest.mark.parametrize("depth", [1, 4])
def test_full_width_reduced_depth_vs_live_oracle(depth):
    """Full-size WIDTHS (CLIP-L d=1024/16 heads/257 tokens, LLaMA h=4096/32 heads/11008, vocab 32000, 224² frames,
    context ≈ 580) at reduced DEPTH, against the oracle run live on the host.  With few layers the random-init
    network cannot amplify rounding noise, so the north-star 1e-2 logit bar applies as written — this is the test that
    pins every full-size kernel shape (BN=256 tiles, K=11008, split-K decode GEMMs, 64-token KV pages, head_dim 128)."""
    from oracle import model as OM
    from oracle import weights as OW
    cfg = _full_width(depth, depth)
    sd = OW.make_state_dict(cfg, 4321)
    model = _model(cfg, 4321)
    ids = [1] + [7 + 3 * i for i in range(30)] + [-200] + [11, 12, 13, 14, 15] + [-200] + [5 + i for i in range(25)]
    frames = OW.synthetic_frames_u8(2, cfg.vision.image_size, 77)
    px = OM.normalize_u8_nhwc(frames)
    n_new = 6
    want, wl = OM.generate_greedy(sd, cfg, ids, px, n_new, policy="bf16", eos_token_id=None, return_logits=True)
    # reproducibility floor of the checker: the same oracle with a different matmul thread count (fp32 accumulation
    # order only).  Random-init full-width layers amplify last-bit differences (≈0.6 % at depth 1, ≈1.1 % at depth 4
    # between two CPU runs), so the bar is the north-star 1e-2 or 2.5× that floor, whichever is larger.
    nt = torch.get_num_threads()
    torch.set_num_threads(3 if nt != 3 else 2)
    _, wl_b = OM.generate_greedy(sd, cfg, ids, px, 1, policy="bf16", eos_token_id=None, return_logits=True)
    torch.set_num_threads(nt)
    floor = ((wl_b[0] - wl[0]).abs().max() / wl[0].abs().max()).item()
    bar = max(LOGIT_RTOL, 2.5 * floor)
    _, wl32 = OM.generate_greedy(sd, cfg, ids, px, 1, policy="fp32", eos_token_id=None, return_logits=True)
    feats = model.encode_images(frames_u8=frames.to(DEV)).float().cpu()
    ref_feats = OM.encode_images(sd, cfg, px, "bf16")
    e_feat = ((feats - ref_feats).abs().max() / ref_feats.abs().max()).item()
    got, gl = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    e32 = ((gl[0, 0].float().cpu() - wl32[0]).abs().max() / wl32[0].abs().max()).item()
    top2 = wl.topk(2, -1).values
    n = compare_tokens(got[0], want, (top2[:, 0] - top2[:, 1]).numpy(), wl.abs().amax(-1).numpy(), f"depth {depth}")
    errs = [((gl[0, s].float().cpu() - wl[s]).abs().max() / wl[s].abs().max()).item() for s in range(n)]
    print(f"depth {depth}: projector-output rel err {e_feat:.2e}; step logits rel err vs bf16-policy oracle {['%.2e' % e for e in errs]}; "
          f"step-0 vs fp32 oracle {e32:.2e}; oracle reproducibility floor {floor:.2e}; tokens verified {n}/{n_new}")
    assert e_feat <= LOGIT_RTOL
    assert n >= 1 and all(e <= bar for e in errs)
    assert e32 <= 3 * bar                               # includes the cost of bf16 storage itself
    del model
    torch.cuda.empty_cache()


def test_graph_replay_equals_eager(tiny):
    from oracle import weights as OW
    cfg, model = tiny
    ids = [[1, 4, -200, 9, 10], [1, -200, -200, 7]]
    frames = [OW.synthetic_frames_u8(1, cfg.vision.image_size, 5), OW.synthetic_frames_u8(2, cfg.vision.image_size, 6)]
    model.use_graph = True
    a = model.generate_batch(ids, frames_u8=frames, max_new_tokens=20)
    model.use_graph = False
    b = model.generate_batch(ids, frames_u8=frames, max_new_tokens=20)
    model.use_graph = True
    assert a == b                  # same kernels, same order: bit-identical
    model.set_pdl(False)           # programmatic dependent launch only changes when kernels start, not what they compute
    c = model.generate_batch(ids, frames_u8=frames, max_new_tokens=20)
    model.use_graph = False
    d = model.generate_batch(ids, frames_u8=frames, max_new_tokens=20)
    model.use_graph = True
    model.set_pdl(True)
    assert a == c == d


def test_truncation_and_errors(tiny):
    from oracle import weights as OW
    cfg, model = tiny
    frames = [OW.synthetic_frames_u8(1, cfg.vision.image_size, 5)]
    with pytest.raises(IndexError):
        model.generate_batch([[1, -200, -200]], frames_u8=frames, max_new_tokens=2)      # more <image> than images
    with pytest.raises(ValueError):
        model.generate_batch([[1, 5, cfg.llama.vocab_size]], frames_u8=frames, max_new_tokens=2)
    # tokenizer_model_max_length truncation of the spliced sequence (llava_arch.py:296-299)
    srcs, lens = model.plan_splice([[1, -200, 7, 8]], [1])
    assert lens[0] == 1 + cfg.tokens_per_image + 2
    model.cfg.tokenizer_model_max_length = 10
    try:
        srcs, lens = model.plan_splice([[1, -200, 7, 8]], [1])
        assert lens[0] == 10 and srcs[0][0] == 1 and srcs[0][1] == -1 and srcs[0][9] == -9
    finally:
        model.cfg.tokenizer_model_max_length = None


def test_reference_api_drop_in(tmp_path):
    """README.md:112-125 usage through the videollava import shim, on PNG files."""
    from PIL import Image

    from videollava.eval.eval import load_model
    from videollava.eval.inference import run_inference_batch, run_inference_single
    tokenizer, model, processor = load_model("teochat-synthetic-tiny?seed=1234", None, device=DEV)
    rng = np.random.RandomState(0)
    paths = []
    for i, size in enumerate([(56, 56), (80, 64)]):
        p = str(tmp_path / f"im{i}.png")
        Image.fromarray(rng.randint(0, 256, (size[0], size[1], 3), dtype=np.uint8)).save(p)
        paths.append(p)
    inp = "This is a sequence of images captured at times: <video> What changed?"
    a = run_inference_single(model, processor, tokenizer, inp, paths, timestamps=["2021-03-01", "2020-01-01"], temperature=0,
                             max_new_tokens=8)
    assert isinstance(a, str) and "</s>" not in a
    b = run_inference_batch(model, processor, tokenizer, [inp], [paths], timestamps_list=[["2021-03-01", "2020-01-01"]],
                            max_new_tokens=8)
    assert b == [a]
    with pytest.raises(ValueError):
        run_inference_single(model, processor, tokenizer, inp, paths, prompt_strategy="bogus", temperature=0)
    # the reference's default decode mode (do_sample=True, temperature=0.2) runs on the device
    c = run_inference_single(model, processor, tokenizer, inp, paths, max_new_tokens=8)
    assert isinstance(c, str)
    d1 = model.generate_batch([[1, -200, 5]], pixel_values=[torch.zeros(1, 3, 56, 56)], max_new_tokens=12, temperature=1.5, seed=7)
    d2 = model.generate_batch([[1, -200, 5]], pixel_values=[torch.zeros(1, 3, 56, 56)], max_new_tokens=12, temperature=1.5, seed=7)
    d3 = model.generate_batch([[1, -200, 5]], pixel_values=[torch.zeros(1, 3, 56, 56)], max_new_tokens=12, temperature=1.5, seed=8)
    assert d1 == d2 and d1 != d3            # graph-replayed sampling is reproducible per seed


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "config1_full.npz")), reason="full-size fixture not generated")
def test_config1_full_size_vs_golden():
    """BASELINE.json configs[0]: 2 frames 224×224, random-init CLIP-L + LLaMA-2-7B, greedy 16 tokens."""
    from oracle import weights as OW
    z = np.load(os.path.join(GOLDEN, "config1_full.npz"))
    cfg = TeoConfig.full()
    model = _model(cfg, int(z["seed"]))
    nf, fs = z["frames_0"].tolist()
    frames = OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs)
    ids = z["ids_0"].tolist()
    outs, logits = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=int(z["max_new"]), return_logits=True)
    l0 = logits[0, 0, ::int(z["logit_stride"])].float().cpu().numpy()
    ref16, ref32 = z["logits0_bf16_0"], z["logits0_fp32_0"]
    err16 = np.abs(l0 - ref16).max() / np.abs(ref16).max()
    err32 = np.abs(l0 - ref32).max() / np.abs(ref32).max()
    gap = np.abs(ref16 - ref32).max() / np.abs(ref32).max()
    print(f"config1 step-0 logits: rel err vs bf16-policy oracle {err16:.3e}, vs fp32 oracle {err32:.3e} "
          f"(oracle bf16-vs-fp32 gap {gap:.3e})")
    # At full size the random-init network amplifies last-bit differences: the SAME oracle run with a different
    # matmul thread count (accumulation order only) moves these logits by `noise` (≈3 %, measured when the fixture
    # was generated, tests/golden/make_golden.py).  No implementation can agree with the fixture tighter than that
    # floor, so the end-to-end bound is the north-star 1e-2 plus twice the floor; the 1e-2 bar itself is enforced
    # where it is meaningful — per layer on identical inputs (test_full_width_reduced_depth_vs_live_oracle) and end to end on the tiny config.
    noise = float(z["selfnoise_bf16_0"])
    print(f"oracle reproducibility floor {noise:.3e}")
    assert err16 <= LOGIT_RTOL + 2 * noise
    assert err32 <= LOGIT_RTOL + 2 * max(noise, gap)
    n = compare_tokens(outs[0], z["tokens_bf16_0"].tolist(), z["margin_bf16_0"], z["absmax_bf16_0"], "config1")
    print("config1 verified greedy steps:", n, "tokens", outs[0])
    # the first id must agree only if the oracle's own top-2 margin there exceeds what two runs of the oracle differ by
    # (fixture: margin 1.7 % of max |logit| vs a 3.3 % floor — a coin flip, and it has flipped between kernel revisions)
    if z["margin_bf16_0"][0] > 2 * noise * z["absmax_bf16_0"][0]:
        assert n >= 1
    del model
    torch.cuda.empty_cache()


def test_load_model_from_hf_directory(tmp_path, tiny):
    """load_model on an HF-format directory (sharded safetensors + config.json) gives the same model as the
    synthetic initialiser with the same seed (the oracle's checkpoint is that initialiser, bit for bit)."""
    import json

    from safetensors.torch import save_file

    from oracle import weights as OW
    from videollava.eval.eval import load_model
    cfg, ref_model = tiny
    sd = OW.make_state_dict(cfg, 1234, dtype=torch.bfloat16)
    d = str(tmp_path / "teochat-tiny-hf")
    os.makedirs(d)
    l, v = cfg.llama, cfg.vision
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump({"hidden_size": l.hidden_size, "intermediate_size": l.intermediate_size, "num_hidden_layers": l.num_hidden_layers,
                   "num_attention_heads": l.num_attention_heads, "vocab_size": l.vocab_size, "rms_norm_eps": l.rms_norm_eps,
                   "max_position_embeddings": l.max_position_embeddings, "mm_projector_type": "mlp2x_gelu", "mm_vision_select_layer": -2,
                   "vision_config": {"hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size,
                                     "num_hidden_layers": v.num_hidden_layers, "num_attention_heads": v.num_attention_heads,
                                     "image_size": v.image_size, "patch_size": v.patch_size}}, f)
    keys = sorted(sd)
    save_file({k: sd[k].contiguous() for k in keys[::2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[1::2]}, os.path.join(d, "model-00002-of-00002.safetensors"))
    tokenizer, model, processor = load_model(d, None, device=DEV)
    for k, t in ref_model.w.t.items():
        assert torch.equal(t, model.w.t[k]), k
    assert processor.crop_size["height"] == v.image_size


def test_cuda_path_vs_reference_code_golden(tmp_path):
    """The CUDA path against outputs of the REFERENCE's own code (tests/golden/reference_path.*: the reference's
    run_inference_single → torchvision transform → CLIPVisionTransformer → feature_select → mlp2x_gelu projector →
    prepare_inputs_labels_for_multimodal, then HF LLaMA greedy; small widths, real 224²/patch-14 geometry).
    Same inputs (the reference's fp16-rounded pixel values), north-star bar: projector output and logits within 1e-2,
    greedy ids equal up to a reference near-tie; then the same examples once more from PNG files through the drop-in
    run_inference_single (GPU preprocessing) — the decoded string equals the reference's."""
    from PIL import Image

    import refgolden
    from teochat_b200.eval.inference import run_inference_single
    from teochat_b200.processor import TeoImageProcessor
    from teochat_b200.tokenizer import StubTokenizer
    cfg, meta, arrays = refgolden.load()
    model = _model(cfg, meta["seed"])
    st, n_new = meta["stride"], meta["max_new"]
    for ci, case in enumerate(meta["cases"]):
        px = torch.from_numpy(arrays[f"pixel_values_f16_{ci}"]).float()
        ids = arrays[f"input_ids_{ci}"].tolist()
        proj = model.encode_images(pixel_values=px.to(DEV)).float().cpu()
        want = torch.from_numpy(arrays[f"projected_{ci}"])
        e_proj = ((proj.flatten()[::st] - want).abs().max() / want.abs().max()).item()
        got, gl = model.generate_batch([ids], pixel_values=[px], max_new_tokens=n_new, return_logits=True)
        wl = torch.from_numpy(arrays[f"logits_{ci}"])
        n = compare_tokens(got[0], arrays[f"tokens_{ci}"].tolist(), arrays[f"margin_{ci}"], wl.abs().amax(-1).numpy(), f"reference case {ci}")
        errs = [((gl[0, s].float().cpu() - wl[s]).abs().max() / wl[s].abs().max()).item() for s in range(n)]
        print(f"reference case {ci}: projector rel err {e_proj:.2e}; logits rel err {['%.2e' % e for e in errs]}; tokens verified {n}/{n_new}")
        assert e_proj <= LOGIT_RTOL and n >= 1 and all(e <= LOGIT_RTOL for e in errs)
        # drop-in API on files: raw u8 → teo_resize_crop_normalize_u8 → … → string
        paths = []
        imgs = [np.random.RandomState(s).randint(0, 256, (h, w, 3), dtype=np.uint8) for h, w, s in case["images"]]
        for k, im in enumerate(imgs):
            p = str(tmp_path / f"c{ci}_{k}.png")
            Image.fromarray(im).save(p)
            paths.append(p)
        out = run_inference_single(model, TeoImageProcessor(224), StubTokenizer(cfg.llama.vocab_size), case["inp"], paths,
                                   timestamps=case["timestamps"], prompt_strategy=case["prompt_strategy"],
                                   chronological_prefix=case["chronological_prefix"], temperature=0, max_new_tokens=n_new)
        if n == n_new:
            assert out == case["output"], (out, case["output"])
    del model
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ BASELINE.json configs at FULL size: size-independent properties
@pytest.fixture(scope="module")
def full():
    cfg = TeoConfig.full()
    model = _model(cfg, 1234)
    yield cfg, model
    del model
    torch.cuda.empty_cache()


def _bench_prompt(cfg, n_frames):
    from teochat_b200.eval.inference import build_prompt
    from teochat_b200.mm_utils import tokenizer_image_token
    from teochat_b200.tokenizer import StubTokenizer
    prompt, _, _ = build_prompt("This is a sequence of images captured at times: <video> What objects or changes can you see across the images?",
                                ["f"] * n_frames)
    return tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size))


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def test_full_size_long_context_properties(full):
    """BASELINE configs[4] per-GPU shape (T=16 frames, 2 sequences per GPU, context ≈ 4.2k — beyond LLaMA-2's 4096
    positions, so the RoPE table past max_position_embeddings and 67-page block tables are exercised) at full size.
    No CPU oracle finishes this in seconds, so the checks are size-independent properties:
      * a sequence duplicated inside the batch gives bit-identical logits and ids (rows are independent, reductions
        are in a fixed order);
      * CUDA-graph replay with PDL == eager launches (ids);
      * KV-cache consistency: the logits of decode step k equal the step-0 logits of a fresh PREFILL over
        prompt + the k generated ids (different kernels for every op: tcgen05 flash vs paged mma decode attention,
        tiled vs swap-AB stream-K GEMMs, fused vs unfused glue) within the bf16 amplification floor of the random-init
        7B network (≈ 3 %); a wrong position, page or mask gives O(1)."""
    from oracle import weights as OW
    cfg, model = full
    T, n_new, k = 16, 24, 20
    ids = _bench_prompt(cfg, T)
    fa, fb = OW.synthetic_frames_u8(T, 224, 41), OW.synthetic_frames_u8(T, 224, 42)
    S0 = len(ids) - T + T * cfg.tokens_per_image
    assert S0 > cfg.llama.max_position_embeddings
    outs, lg = model.generate_batch([ids, ids, ids], frames_u8=[fa, fb, fa], max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    assert all(len(o) == n_new and all(0 <= t < cfg.llama.vocab_size for t in o) for o in outs)
    assert outs[0] == outs[2] and torch.equal(lg[0], lg[2])
    assert outs[0] != outs[1]                                   # different frames → different continuation
    graph = model.generate_batch([ids, ids, ids], frames_u8=[fa, fb, fa], max_new_tokens=n_new, eos_token_id=-1)
    assert graph == outs
    forced, lf = model.generate_batch([ids + outs[1][:k]], frames_u8=[fb], max_new_tokens=1, eos_token_id=-1, return_logits=True)
    err = _rel(lf[0, 0], lg[1, k])
    print(f"long context S0={S0}: decode step {k} vs re-prefill logits rel err {err:.3e}; same next id: {forced[0][0] == outs[1][k]}")
    assert err <= 8e-2


def test_full_size_single_image_batch64_properties(full):
    """BASELINE configs[1] shape (T=1, bs=64, 128 new tokens, GeoChat-style single image) at full size: every sequence
    produces 128 ids; a duplicated sample is bit-identical; a sample's first logits do not depend on its batch
    (bs=64 → BN=64 swap-AB decode tiles, bs=2 → BN=32) beyond the amplification floor."""
    from oracle import weights as OW
    cfg, model = full
    B, n_new = 64, 128
    ids = _bench_prompt(cfg, 1)
    frames = [OW.synthetic_frames_u8(1, 224, 100 + b) for b in range(B)]
    frames[63] = frames[5]
    outs = model.generate_batch([ids] * B, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1)
    assert len(outs) == B and all(len(o) == n_new for o in outs)
    assert outs[63] == outs[5] and outs[5] != outs[6]
    _, l64 = model.generate_batch([ids] * B, frames_u8=frames, max_new_tokens=2, eos_token_id=-1, return_logits=True)
    _, l2 = model.generate_batch([ids] * 2, frames_u8=frames[5:7], max_new_tokens=2, eos_token_id=-1, return_logits=True)
    e0, e1 = _rel(l64[5, 0], l2[0, 0]), _rel(l64[5, 1], l2[0, 1])
    print(f"bs=64 vs bs=2, sample 5: prefill logits rel err {e0:.3e}, first decode step {e1:.3e}")
    assert e0 <= 8e-2 and e1 <= 8e-2


@pytest.mark.parametrize("case", ["tiny_b5", "tiny_b40", "tiny_b100", "fullwidth_depth2_b3"])
def test_decode_chain_bit_identical_to_kernel_per_gemm(case):
    """The persistent decode chain kernel (decode_chain.cu: o_proj → norm → gate/up → SwiGLU → down → norm → next qkv / lm_head in
    one launch per layer, grid barriers between the phases) against the one-kernel-per-GEMM sequence it replaces: same MMA
    order, same reduction order → every step's logits and all ids BIT-identical, eager and graph-replayed, at the three
    batch tile widths (BN = 32 / 64 / 128) and at full LLaMA-2-7B widths."""
    from oracle import weights as OW
    if case.startswith("tiny"):
        cfg, B, n_new = TeoConfig.tiny(), int(case.split("_b")[1]), 12
    else:
        cfg, B, n_new = _full_width(2, 1), 3, 8
    model = _model(cfg, 99)
    ids = [[1, 17, -200, 5, 6, 30 + b % 7] + ([-200, 9] if b % 3 == 0 else []) for b in range(B)]
    frames = [OW.synthetic_frames_u8(2 if b % 3 == 0 else 1, cfg.vision.image_size, 200 + b % 11) for b in range(B)]
    model.set_decode_chain(True)
    ids_c, lg_c = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    graph_c = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1)
    launches_c = model.decode_step_launches
    model.set_decode_chain(False)
    ids_k, lg_k = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    graph_k = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1)
    launches_k = model.decode_step_launches
    _, lg_k2 = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    model.set_decode_chain(True)
    _, lg_c2 = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
    model.set_decode_chain(False)
    assert torch.isfinite(lg_c).all()
    # each path must first reproduce ITSELF run to run (fixed reduction orders everywhere), then the two must agree
    assert torch.equal(lg_k, lg_k2), f"per-GEMM path not reproducible: max |diff| {(lg_k - lg_k2).abs().max().item():.3e}"
    assert torch.equal(lg_c, lg_c2), f"chain path not reproducible: max |diff| {(lg_c - lg_c2).abs().max().item():.3e}"
    assert torch.equal(lg_c, lg_k), f"max |diff| {(lg_c - lg_k).abs().max().item():.3e}"
    assert ids_c == ids_k == graph_c == graph_k
    print(f"{case}: kernels per decode step {launches_c} (chain) vs {launches_k} (per GEMM)")
    assert launches_c < launches_k
    del model
    torch.cuda.empty_cache()


def test_stream_k_fused_swiglu_reduction_is_bit_identical(tmp_path):
    """TEO_SK_FUSE=1 (gemm.cu, SkFuse): the decode gate/up GEMM reduces its own stream-K partials and applies SwiGLU — the CTA holding slot 0
    of a weight tile waits for the other slots' arrival flags and runs the glue kernel's reduction code on that tile — against the
    stand-alone reduce_swiglu kernel: logits and ids BIT-identical, eager and graph-replayed (the switch is read once per process, so
    each setting runs in its own process: tools/sk_fuse_check.py).  Optional path, measured 0.3-0.6 % slower (profiles/r02_sk_fuse.txt)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = []
    for f in ("0", "1"):
        out = str(tmp_path / f"skf{f}.pt")
        env = dict(os.environ, TEO_SK_FUSE=f)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "sk_fuse_check.py"), "dump", out], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        files.append(out)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sk_fuse_check.py"), "cmp", *files], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.count("BIT-IDENTICAL") == 4, r.stdout[-2000:] + r.stderr[-2000:]


def test_vit_folded_layernorm_path_matches_default(monkeypatch):
    """The optional folded-LayerNorm ViT (TEO_VIT_LN_FOLD=1: teo_gemm_bf16_ex with gain-scaled weights, row statistics from the
    out-proj / fc2 epilogues, no LayerNorm kernels) against the default path and the fp32 oracle at full CLIP-L width."""
    from oracle import model as OM
    from oracle import weights as OW
    cfg = _full_width(1, 2)
    frames = OW.synthetic_frames_u8(3, cfg.vision.image_size, 55)
    base = _model(cfg, 77)
    l0 = base.launch_count()
    ref = base.encode_images(frames_u8=frames.to(DEV)).float().cpu()
    n_default = base.launch_count() - l0
    monkeypatch.setenv("TEO_VIT_LN_FOLD", "1")
    folded = _model(cfg, 77)
    assert folded.w.ln_folded and "vit.0.qkv_wf" in folded.w.t
    l0 = folded.launch_count()
    got = folded.encode_images(frames_u8=frames.to(DEV)).float().cpu()
    n_folded = folded.launch_count() - l0
    sd = OW.make_state_dict(cfg, 77, prefixes=[OM.VIT, "model.mm_projector"])
    want = OM.encode_images(sd, cfg, OM.normalize_u8_nhwc(frames), "fp32")
    e_def, e_fold, e_pair = _rel(ref, want), _rel(got, want), _rel(got, ref)
    print(f"ViT+projector vs fp32 oracle: default {e_def:.2e}, folded LayerNorm {e_fold:.2e}; folded vs default {e_pair:.2e}; "
          f"launches {n_default} -> {n_folded}")
    assert e_fold <= 1e-2 and e_pair <= 1e-2 and n_folded < n_default
    del base, folded
    torch.cuda.empty_cache()
