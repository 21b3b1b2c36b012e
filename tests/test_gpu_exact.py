"""GPU parity tests of the EXACT mode (include/teochat_b200.h "Exact mode"; csrc/exact.cu, model_exact.cu): fp32 activations,
residual stream and KV pages, split-bf16 tensor-core GEMMs with fp32 accumulation — the north-star clause "bit-exact token
ids under greedy decode with fp32 accumulation".  Checked against the fp32 oracle: committed golden vectors (tiny config and
BASELINE.json configs[0] at FULL size and depth) and the live oracle at the benchmark contexts (ctx ≈ 2130 ragged bs=4,
ctx ≈ 4240) at full width.  Bars: logits ≤ 1e-2 (north star) — and the measured 1e-4-level agreement is asserted at 1e-3 so a
regression to bf16-level error cannot hide; greedy ids equal, unconditionally.
"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err, stream
from teochat_b200 import lib as L
from teochat_b200.config import TeoConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NORTH_STAR = 1e-2
EXACT_BAR = 1e-3


def _model(cfg, seed=1234, precision="exact"):
    from teochat_b200.engine import TeoModel
    from teochat_b200.weights import TeoWeights
    return TeoModel(cfg, TeoWeights.from_synthetic(cfg, seed, DEV), DEV, precision=precision)


@pytest.mark.parametrize("M,N,K,blocked", [(5, 512, 256, False), (300, 1024, 1024, True), (64, 4096, 4096, True), (257 * 2, 384, 128, False),
                                           (1000, 128, 192, False), (33, 32000, 4096, True)])
def test_gemm_bf16x3_is_an_fp32_linear(teo, M, N, K, blocked):
    """teo_split_f32_bf16x3 + teo_gemm_bf16x3 against a float64 nn.Linear on the same fp32 inputs / bf16 weights."""
    lib, h = teo
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))          # rows of very different scale
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = (torch.randn(N, generator=g) * 0.1).to(torch.bfloat16)
    res = torch.randn(M, N, generator=g)
    want = (x.double() @ w.double().T + bias.double() + res.double()).float()
    xd, wd, bd, out = x.to(DEV), w.to(DEV), bias.to(DEV), res.to(DEV).clone()
    planes = torch.empty(M, 3 * K, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_split_f32_bf16x3(xd.data_ptr(), planes.data_ptr(), M, K, stream()))
    p3 = planes.float().view(M, 3, K)
    assert torch.equal(p3.sum(1), xd)                                   # hi + mid + lo == x, exactly
    if blocked:
        wb = torch.empty_like(wd)
        L.check(lib.teo_weight_to_blocked(wd.data_ptr(), wb.data_ptr(), N, K, stream()))
        wd = wb
    ws = torch.empty(max(16, lib.teo_gemm_workspace_bytes(M, N, 3 * K)), dtype=torch.uint8, device=DEV)
    L.check(lib.teo_gemm_bf16x3(h, planes.data_ptr(), wd.data_ptr(), int(blocked), out.data_ptr(), M, N, K, bd.data_ptr(), out.data_ptr(),
                                ws.data_ptr(), ws.numel(), stream()), "teo_gemm_bf16x3")
    err = rel_err(out.cpu(), want)
    print(f"bf16x3 GEMM {M}x{N}x{K}: rel err vs float64 {err:.2e}")
    # the inputs are exact; what is left is the fp32 accumulation inside the tensor core over 3K terms (one TMEM accumulator chain
    # per stream-K slice: measured 2e-7 at K=256 … 2e-5 at K=4096 with few slices), far below the bf16 path's 4e-3 per GEMM
    assert err <= 5e-5


def test_exact_tiny_vs_golden_fp32():
    """Every sample of the tiny fixture: all 24 greedy ids equal the fp32 oracle's, step-0 logits within 1e-4."""
    from oracle import weights as OW
    cfg = TeoConfig.tiny()
    model = _model(cfg)
    z = np.load(os.path.join(GOLDEN, "tiny_generate.npz"))
    ids, frames = [], []
    for i in range(int(z["n_samples"])):
        ids.append(z[f"ids_{i}"].tolist())
        nf, fs = z[f"frames_{i}"].tolist()
        frames.append(OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs))
    outs, logits = model.generate_batch(ids, frames_u8=frames, max_new_tokens=int(z["max_new"]), return_logits=True)
    for i, out in enumerate(outs):
        ref = z[f"logits0_fp32_{i}"]
        err = np.abs(logits[i, 0, ::int(z["logit_stride"])].cpu().numpy() - ref).max() / np.abs(ref).max()
        print(f"tiny sample {i}: step-0 logits rel err vs fp32 oracle {err:.2e}")
        assert err <= 1e-4
        assert out == z[f"tokens_fp32_{i}"].tolist(), f"sample {i}"
    graph = model.generate_batch(ids, frames_u8=frames, max_new_tokens=int(z["max_new"]))       # CUDA-graph replay of the exact step
    assert graph == outs


def test_exact_tiny_vs_live_oracle_every_step():
    from oracle import model as OM
    from oracle import weights as OW
    cfg = TeoConfig.tiny()
    model = _model(cfg, 777)
    sd = OW.make_state_dict(cfg, 777)
    ids = [1, 17, 99, -200, 5, 6, -200, 300, 301, 302, -200, 9]
    frames = OW.synthetic_frames_u8(3, cfg.vision.image_size, 31)
    px = OM.normalize_u8_nhwc(frames)
    feats = model.encode_images(frames_u8=frames.to(DEV)).cpu()
    assert feats.dtype == torch.float32
    assert rel_err(feats, OM.encode_images(sd, cfg, px, "fp32")) <= 1e-5
    want, wl = OM.generate_greedy(sd, cfg, ids, px, 20, policy="fp32", eos_token_id=None, return_logits=True)
    got, gl = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=20, eos_token_id=-1, return_logits=True)
    errs = [rel_err(gl[0, s].cpu(), wl[s]) for s in range(20)]
    print("exact tiny, per-step logits rel err vs fp32 oracle:", ["%.1e" % e for e in errs])
    assert max(errs) <= 1e-4
    assert got[0] == want


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "config1_full.npz")), reason="full-size fixture not generated")
def test_exact_config1_full_size_all_ids():
    """BASELINE.json configs[0] at FULL size and depth (2 frames 224², CLIP-L 23 layers, LLaMA-2-7B 32 layers, context 580,
    greedy 16 tokens): step-0 logits within the north-star 1e-2 of the fp32 oracle (measured ~1e-5, asserted at 1e-3) and
    16 of 16 greedy ids equal — unconditionally (the fp32 oracle's smallest top-2 margin is 1.1e-3 of max |logit|)."""
    from oracle import weights as OW
    z = np.load(os.path.join(GOLDEN, "config1_full.npz"))
    cfg = TeoConfig.full()
    model = _model(cfg, int(z["seed"]))
    nf, fs = z["frames_0"].tolist()
    frames = OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs)
    ids = z["ids_0"].tolist()
    outs, logits = model.generate_batch([ids], frames_u8=[frames], max_new_tokens=int(z["max_new"]), return_logits=True)
    ref32 = z["logits0_fp32_0"]
    err = np.abs(logits[0, 0, ::int(z["logit_stride"])].cpu().numpy() - ref32).max() / np.abs(ref32).max()
    want = z["tokens_fp32_0"].tolist()
    n_eq = sum(int(a == b) for a, b in zip(outs[0], want))
    print(f"config1 exact mode: step-0 logits rel err vs fp32 oracle {err:.3e}; greedy ids equal {n_eq}/{len(want)}; ids {outs[0]}")
    assert err <= NORTH_STAR and err <= EXACT_BAR
    assert outs[0] == want
    del model
    torch.cuda.empty_cache()


def _full_width(depth):
    cfg = TeoConfig.full()
    cfg.llama.num_hidden_layers = depth
    cfg.vision.num_hidden_layers = depth + 1
    return cfg


def _prompt(cfg, n_frames, extra):
    from teochat_b200.eval.inference import build_prompt
    from teochat_b200.mm_utils import tokenizer_image_token
    from teochat_b200.tokenizer import StubTokenizer
    prompt, _, _ = build_prompt("This is a sequence of images captured at times: <video> What objects or changes can you see across the images?"
                                + " and" * extra, ["f"] * n_frames)
    return tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size))


@pytest.mark.parametrize("shape", ["ctx2130_bs4_ragged", "ctx4240"])
def test_benchmark_contexts_vs_live_oracle(shape):
    """The BENCHMARK contexts at full width (depth 1), oracle-checked: BASELINE configs[2] (T=8 → context ≈ 2130) as a RAGGED
    batch of 4 (7/8/8/6 frames, different prompt lengths: exercises cu_seqlens, per-sequence page tables and positions) and
    configs[4] (T=16 → context ≈ 4240 > LLaMA-2's 4096 positions).  Prefill logits AND two decode steps per sequence:
      exact mode vs fp32 oracle  ≤ 1e-3 (north star 1e-2), greedy ids equal;
      bf16 mode  vs bf16-policy oracle ≤ 1e-2 (the north-star logit bar on the product path at the benchmark context)."""
    from oracle import model as OM
    from oracle import weights as OW
    cfg = _full_width(1)
    seed = 2468
    sd = OW.make_state_dict(cfg, seed)
    if shape == "ctx4240":
        specs = [(16, 0)]
    else:
        specs = [(7, 0), (8, 3), (8, 11), (6, 40)]
    ids = [_prompt(cfg, t, extra) for t, extra in specs]
    frames = [OW.synthetic_frames_u8(t, cfg.vision.image_size, 500 + i) for i, (t, _) in enumerate(specs)]
    ctx = [len(i) - t + t * cfg.tokens_per_image for i, (t, _) in zip(ids, specs)]
    n_new = 3
    want = {}
    for pol in ("fp32", "bf16"):
        want[pol] = [OM.generate_greedy(sd, cfg, ids[i], OM.normalize_u8_nhwc(frames[i]), n_new, policy=pol, eos_token_id=None, return_logits=True)
                     for i in range(len(ids))]
    # reproducibility floor of the bf16-policy checker itself (same oracle, another matmul thread count: accumulation order only)
    nt = torch.get_num_threads()
    torch.set_num_threads(3 if nt != 3 else 2)
    _, wl_b = OM.generate_greedy(sd, cfg, ids[0], OM.normalize_u8_nhwc(frames[0]), 1, policy="bf16", eos_token_id=None, return_logits=True)
    torch.set_num_threads(nt)
    floor = rel_err(wl_b[0], want["bf16"][0][1][0])
    print(f"{shape}: bf16-policy oracle reproducibility floor {floor:.2e}")
    for precision, pol, bar in (("exact", "fp32", EXACT_BAR), ("bf16", "bf16", max(NORTH_STAR, 2.5 * floor))):
        model = _model(cfg, seed, precision)
        got, gl = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
        for i in range(len(ids)):
            wt, wl = want[pol][i]
            errs = [rel_err(gl[i, s].cpu(), wl[s]) for s in range(n_new)]
            print(f"{shape} seq {i} (context {ctx[i]}): {precision} mode vs {pol} oracle, logits rel err per step {['%.1e' % e for e in errs]}; "
                  f"ids {got[i]} vs {wt}")
            if precision == "exact":
                assert max(errs) <= bar and got[i] == wt
            else:
                # step 0 is the prefill; later steps are comparable while the fed-back ids agree
                assert errs[0] <= bar
                for s in range(1, n_new):
                    if got[i][:s] == wt[:s]:
                        assert errs[s] <= bar
        del model
        torch.cuda.empty_cache()
