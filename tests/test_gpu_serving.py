"""GPU tests of the serving-side behaviour around the kernels: retirement of finished sequences, the drop-in run_inference
loop on a TEOChatlas-shaped dataset, and the data-parallel run (2 ranks over NCCL == 1 rank)."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from teochat_b200.config import TeoConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(cfg, seed=1234):
    from teochat_b200.engine import TeoModel
    from teochat_b200.weights import TeoWeights
    return TeoModel(cfg, TeoWeights.from_synthetic(cfg, seed, DEV), DEV)


def test_finished_sequences_are_retired():
    """Each sample stops at its own </s> (inference.py:57-72).  Rows that finish leave the batch at the next 32-step sync
    (state compaction + a graph for the smaller batch, pages returned through teo_kv_free); survivors continue and produce
    the ids they produce without retirement."""
    from oracle import weights as OW
    cfg = TeoConfig.tiny()
    model = _model(cfg)
    ids = [[1, 17, -200, 5, 6, 300 + b] for b in range(8)]
    frames = [OW.synthetic_frames_u8(1, cfg.vision.image_size, 60 + (0 if b < 3 else b)) for b in range(8)]
    for b in range(3):
        ids[b] = ids[0]                                   # rows 0-2 are the same example: they finish together
    free = model.generate_batch(ids, frames_u8=frames, max_new_tokens=100, eos_token_id=-1)
    assert free[0] == free[1] == free[2]
    eos = free[0][6]                                      # an id rows 0-2 emit at step <= 6
    model.retire_finished = False
    plain = model.generate_batch(ids, frames_u8=frames, max_new_tokens=100, eos_token_id=eos, time_phases=True)
    t_plain = dict(model.last_timings)
    model.retire_finished = True
    retired = model.generate_batch(ids, frames_u8=frames, max_new_tokens=100, eos_token_id=eos, time_phases=True)
    t_ret = dict(model.last_timings)
    assert all(o[-1] == eos or len(o) == 100 for o in plain)
    assert len(plain[0]) <= 7 and plain[0][-1] == eos
    assert t_plain["final_batch"] == 8 and t_ret["final_batch"] <= 5 and t_ret["retired_pages"] > 0
    assert retired == plain, [(len(a), len(b)) for a, b in zip(retired, plain)]
    # a second call reuses the cached states / graphs of both batch shapes
    assert model.generate_batch(ids, frames_u8=frames, max_new_tokens=100, eos_token_id=eos) == plain


def _fake_dataset(tmp_path, n=5):
    from PIL import Image
    rng = np.random.RandomState(3)
    data = []
    for i in range(n):
        paths = []
        for k in range(1 + i % 3):
            p = str(tmp_path / f"ex{i}_{k}.png")
            Image.fromarray(rng.randint(0, 256, (56 + 8 * (i % 2), 56, 3), dtype=np.uint8)).save(p)
            paths.append(p)
        q = "This is a sequence of images captured at times: <video> What changed?"
        a = "nothing"
        ex = {"video": paths, "timestamp": [f"20{20 - k}-01-0{k + 1}" for k in range(len(paths))], "task": ["change_detection", "qa", "localization"][i % 3],
              "conversations": [{"from": "human", "value": q}, {"from": "gpt", "value": a}]}
        if i == 1:
            ex["polygon"] = "POLYGON ((0 0, 0 10, 10 10, 10 0, 0 0))"
        if i == 2:
            ex["conversations"][0]["value"] = q + " Region [10, 20, 30, 40]."
            ex["conversations"][1]["value"] = "[1, 2, 3, 4] and [5, 6, 7, 8]"
        data.append(ex)
    return data


def test_run_inference_on_a_dataset(tmp_path):
    """run_inference (videollava/eval/inference.py:88-137) end to end on the GPU over a 5-example TEOChatlas-shaped dataset:
    the reference's bs=1 loop and the batched route give the same strings and the same result dicts (keys response,
    ground_truth, task, and — only where the reference adds them — polygon, input_bboxes, output_bboxes)."""
    from videollava.eval.eval import load_model
    from videollava.eval.inference import run_inference, run_inference_single
    tokenizer, model, processor = load_model("teochat-synthetic-tiny?seed=1234", None, device=DEV)
    data = _fake_dataset(tmp_path)
    args = dict(prompt_strategy="interleave", chronological_prefix=True, conv_mode="v1", temperature=0, max_new_tokens=12)
    one = run_inference(data, model, tokenizer, processor, **args)
    batched = run_inference(data, model, tokenizer, processor, batch_size=3, **args)
    assert one == batched and len(one) == 5
    for ex, out in zip(data, one):
        assert out["ground_truth"] == ex["conversations"][1]["value"] and out["task"] == ex["task"]
        assert isinstance(out["response"], str) and "</s>" not in out["response"]
        single = run_inference_single(model, processor, tokenizer, ex["conversations"][0]["value"], ex["video"], timestamps=ex["timestamp"],
                                      temperature=0, max_new_tokens=12)
        assert out["response"] == single
    assert set(one[0]) == {"response", "ground_truth", "task"}
    assert one[1]["polygon"] == data[1]["polygon"] and "polygon" not in one[0]
    assert one[2]["input_bboxes"] == [[10, 20, 30, 40]] and one[2]["output_bboxes"] == [[1, 2, 3, 4], [5, 6, 7, 8]]
    json.dumps(one)                                          # what eval() writes to the results file


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank(tmp_path):
    """SURVEY.md §4 "distributed": 2 ranks over NCCL, each generating a contiguous shard with its own replica, one
    all_gather_into_tensor at the end — the gathered ids equal the ids of a single-rank run over the same shards, in dataset order."""
    script = tmp_path / "dp.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json
        sys.path.insert(0, {ROOT!r})
        import torch
        from oracle import weights as OW
        from teochat_b200 import dist as TD
        from teochat_b200.config import TeoConfig
        from teochat_b200.engine import TeoModel
        from teochat_b200.weights import TeoWeights
        rank, world, local = TD.init_from_env()
        dev = f"cuda:{{local}}"
        torch.cuda.set_device(local)
        cfg = TeoConfig.tiny()
        model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
        n_total, max_new = 7, 20
        ids = [[1, 17, -200, 5, 6, 300 + i] + ([-200, 9] if i % 2 else []) for i in range(n_total)]
        frames = [OW.synthetic_frames_u8(2 if i % 2 else 1, cfg.vision.image_size, 70 + i) for i in range(n_total)]
        lo, hi = TD.shard_range(n_total, rank, world)
        outs = model.generate_batch(ids[lo:hi], frames_u8=frames[lo:hi], max_new_tokens=max_new, eos_token_id=-1)
        rows = max(h - l for l, h in (TD.shard_range(n_total, r, world) for r in range(world)))
        got = TD.gather_tokens(TD.pack_tokens(outs, rows, max_new, dev), n_total)
        if rank == 0:
            # the single-rank run generates the SAME shards one after the other (a sequence's bits depend on the batch it is decoded in
            # only through the KV-split choice of the attention kernel, so equal batch shapes make the comparison exact)
            want = []
            for r in range(world):
                a, b = TD.shard_range(n_total, r, world)
                want += model.generate_batch(ids[a:b], frames_u8=frames[a:b], max_new_tokens=max_new, eos_token_id=-1)
            assert got == want, (got, want)
            whole = model.generate_batch(ids, frames_u8=frames, max_new_tokens=max_new, eos_token_id=-1)
            assert sum(int(x[:6] == y[:6]) for x, y in zip(whole, got)) == n_total, (whole, got)     # one batch of 7: same ids up to near-ties
            print("OK", torch.distributed.get_backend())
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    """))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK nccl" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
