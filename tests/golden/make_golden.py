"""Generates the committed golden vectors from the CPU oracle (run here, in the build container;
nothing in it can run on the GPU box's -m gpu path).

    python tests/golden/make_golden.py tiny     → tests/golden/tiny_generate.npz
    python tests/golden/make_golden.py full     → tests/golden/config1_full.npz   (BASELINE.json configs[0])

Inputs are fully determined by (seed, frame seed, prompt ids) which are stored in the fixture;
weights come from oracle/hashinit (bit-reproducible).  For every sample the fixture holds the
greedy token ids under both oracle policies, the per-step top-2 logit margin (so a consumer can
tell a real divergence from a near-tie) and a strided slice of the step-0 logits.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import model as OM          # noqa: E402
from oracle import weights as OW        # noqa: E402
from teochat_b200.config import TeoConfig                      # noqa: E402  (shape spec + host glue only)
from teochat_b200.eval.inference import build_prompt           # noqa: E402
from teochat_b200.mm_utils import tokenizer_image_token        # noqa: E402
from teochat_b200.tokenizer import StubTokenizer               # noqa: E402

INSTRUCTION = ("This is a sequence of images captured at times: <video> "
               "What objects or changes can you see across the images?")
LOGIT_STRIDE = 97


def prompt_ids(cfg, n_frames):
    prompt, _, _ = build_prompt(INSTRUCTION, ["f"] * n_frames)
    return tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size))


def run(cfg, seed, samples, max_new, out_path, selfnoise=False):
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    sd = OW.make_state_dict(cfg, seed)
    print(f"weights {time.time() - t0:.1f}s", flush=True)
    rec = {"seed": seed, "max_new": max_new, "n_samples": len(samples), "logit_stride": LOGIT_STRIDE}
    for i, (n_frames, frame_seed) in enumerate(samples):
        ids = prompt_ids(cfg, n_frames)
        frames = OW.synthetic_frames_u8(n_frames, cfg.vision.image_size, frame_seed)
        px = OM.normalize_u8_nhwc(frames)
        rec[f"ids_{i}"] = np.asarray(ids, dtype=np.int64)
        rec[f"frames_{i}"] = np.asarray([n_frames, frame_seed], dtype=np.int64)
        for pol in ("bf16", "fp32"):
            t0 = time.time()
            toks, logits = OM.generate_greedy(sd, cfg, ids, px, max_new, policy=pol, eos_token_id=cfg.llama.eos_token_id,
                                              return_logits=True)
            top2 = logits.topk(2, dim=-1).values
            rec[f"tokens_{pol}_{i}"] = np.asarray(toks, dtype=np.int64)
            rec[f"margin_{pol}_{i}"] = (top2[:, 0] - top2[:, 1]).numpy()
            rec[f"absmax_{pol}_{i}"] = logits.abs().amax(dim=-1).numpy()
            rec[f"logits0_{pol}_{i}"] = logits[0, ::LOGIT_STRIDE].numpy()
            print(f"sample {i} policy {pol}: {time.time() - t0:.1f}s tokens {toks}", flush=True)
    if selfnoise:
        # Reproducibility floor of the oracle itself: rerun the bf16-policy prefill with a different matmul thread
        # count (only the fp32 accumulation order changes).  At full size with random-init weights the logits move by
        # a few percent — the network amplifies last-bit differences — so no implementation can match tighter.
        torch.set_num_threads(3)
        ids = rec["ids_0"].tolist()
        nf, fs = rec["frames_0"].tolist()
        px = OM.normalize_u8_nhwc(OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs))
        _, lg = OM.generate_greedy(sd, cfg, ids, px, 1, policy="bf16", return_logits=True)
        a, b = lg[0, ::LOGIT_STRIDE].numpy(), rec["logits0_bf16_0"]
        rec["selfnoise_bf16_0"] = np.float64(np.abs(a - b).max() / np.abs(b).max())
        print("oracle self-noise (3 vs all threads):", rec["selfnoise_bf16_0"], flush=True)
    np.savez_compressed(out_path, **rec)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    here = os.path.dirname(os.path.abspath(__file__))
    if which == "tiny":
        run(TeoConfig.tiny(), 1234, [(2, 11), (1, 12), (3, 13), (8, 14)], 24, os.path.join(here, "tiny_generate.npz"))
    elif which == "full":
        run(TeoConfig.full(), 1234, [(2, 11)], 16, os.path.join(here, "config1_full.npz"), selfnoise=True)
    else:
        raise SystemExit("usage: make_golden.py tiny|full")
