"""Benchmark of the TEOChat inference hot path (contract in the task statement, §④).

    python bench.py [--gpus N --steps K --warmup W] [--config 1|2|4]     this build on N B200s
    python bench.py --impl reference [...]                              the reference's CPU path on the host cores

One "step" = one full pass of the hot path over one batch of synthetic input: ViT encode of all
frames → projector → splice → ragged prefill → greedy decode of `new` tokens.  The headline workload at every
N is BASELINE.json configs[2] per GPU (T=8 frames, bs=32, 256 new tokens; configs[3] is the same
per-GPU work on 8 GPUs): weak scaling, examples sharded across ranks, one NCCL all-gather of ids
at the end.  `value` = generated tokens/s with the frames already resident in HBM; `e2e` = the
same through the public batched API with frames in pinned HOST memory (H2D of the frames and D2H
of the ids inside the timed region).  At N=1 the line also carries `other_configs`: short measurements of configs[1]
(T=1, bs=64, 128 new tokens) and of the per-GPU shape of configs[4] (T=16, bs=2 per GPU, 512 new tokens, 4k context);
`--config 1|4` makes one of them the headline workload instead (e.g. configs[4] at its true bs=16 under torchrun on 8 GPUs).

CPU legs (`--impl reference`, and `cpu_baseline` of the GPU line): BASELINE.json configs[0] — one 2-frame sequence, context
≈ 580, greedy 16 tokens — end to end through the INSTALLED HF modules the way the reference drives them
(oracle/hf_reference.py), fp32, all usable host cores.  It is a different (smaller) config than the GPU
line and a port, not the reference package: both facts are in the line (`same_config`, `cpu_baseline.kind`).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated_tokens_per_s"
UNIT = "tokens/s"
# CPU legs: BASELINE.json configs[0] — one sequence of T=2 frames (context ≈ 580), greedy 16 tokens
CPU_SAMPLE = (2, 16)
# BASELINE.json configs → (frames T, examples per GPU, new tokens); configs[3] is configs[2] on 8 GPUs, configs[4] is bs=16 over 8 GPUs
CONFIGS = {1: (1, 64, 128), 2: (8, 32, 256), 4: (16, 2, 512)}
INSTRUCTION = ("This is a sequence of images captured at times: <video> "
               "What objects or changes can you see across the images?")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 6 and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_prompt_ids(cfg, n_frames):
    from teochat_b200.eval.inference import build_prompt
    from teochat_b200.mm_utils import tokenizer_image_token
    from teochat_b200.tokenizer import StubTokenizer
    prompt, _, _ = build_prompt(INSTRUCTION, ["f"] * n_frames)
    return tokenizer_image_token(prompt, StubTokenizer(cfg.llama.vocab_size))


# ------------------------------------------------------------------------------------------ CPU legs
def host_cpus() -> int:
    """CPUs this process may really use: the affinity mask, cut to the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def pick_cpu_threads() -> int:
    """Fixed policy: every CPU this process may use (affinity mask ∩ cgroup quota, host_cpus()); TEO_CPU_THREADS overrides."""
    if os.environ.get("TEO_CPU_THREADS"):
        return max(1, int(os.environ["TEO_CPU_THREADS"]))
    return host_cpus()


def cpu_reference_leg(n_frames: int, new_tokens: int, steps: int, warmup: int, seed: int = 1234, budget_s: float = None):
    """BASELINE configs[0] through the installed HF modules (oracle/hf_reference.py; kind "port": the reference package
    cannot be imported) on all usable host cores: fp32, random-init full-size weights, one sequence per step."""
    import torch

    from oracle import hf_reference as HR
    from oracle import weights as OW
    from teochat_b200.config import TeoConfig
    cfg = TeoConfig.full()
    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    sd = OW.make_state_dict(cfg, seed, dtype=torch.float32)
    modules = HR.build_modules(cfg, sd)
    ids = make_prompt_ids(cfg, n_frames)
    frames = OW.synthetic_frames_u8(n_frames, cfg.vision.image_size, 11)
    times, phases = [], []
    it = 0
    while it < warmup + steps:
        t0 = time.perf_counter()
        toks, ph = HR.run_inference_greedy(modules, cfg, sd, ids, frames, new_tokens)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            phases.append(ph)
        if it == 0 and budget_s is not None and (warmup + steps) * dt > budget_s:
            # a slow host: keep the whole run within the budget (>= 1 warm-up when any was asked for, >= 3 timed steps)
            fit = max(3, int(budget_s / dt) - 1)
            if warmup + steps > fit:
                warmup = min(warmup, 1)
                steps = max(3, min(steps, fit - warmup))
        it += 1
    t = sum(times) / len(times)
    mean = lambda k: sum(p[k] for p in phases) / len(phases)
    return {"value": new_tokens / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"BASELINE configs[0]: 1 sequence x T={n_frames} frames, context {phases[0]['context']}, greedy {new_tokens} new tokens, "
                      f"installed HF CLIPVisionModel + LlamaForCausalLM (eager) fp32 driven like videollava.eval.inference; "
                      f"mean of {len(times)} run(s) after {warmup} warm-up",
            "s_per_sample": t, "s_min": min(times), "s_max": max(times), "steps_run": len(times), "warmup_run": warmup,
            "vit_frames_per_s": n_frames / mean("vision_s"), "prefill_tokens_per_s": phases[0]["context"] / mean("prefill_s"),
            "decode_tokens_per_s": (new_tokens - 1) / mean("decode_s")}


def workload_name(T, B, new, S0=None):
    ctx = f", context {S0}" if S0 else ""
    tag = {(1, 64, 128): "BASELINE configs[1]", (8, 32, 256): "BASELINE configs[2]/[3]", (16, 2, 512): "BASELINE configs[4] per-GPU shape"}.get((T, B, new), "custom")
    return f"T={T} frames x bs={B} per GPU{ctx}, {new} new tokens, greedy ({tag})"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, new = CPU_SAMPLE                     # bounded sample per step: BASELINE configs[0] (the full GPU config would take hours on CPU)
    cb = cpu_reference_leg(T, new, args.steps, args.warmup, budget_s=float(os.environ.get("TEO_CPU_BUDGET_S", "420")))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps_run"],
            "warmup": cb["warmup_run"], "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": cb["s_per_sample"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "kind": cb["kind"], "same_config": False,
            "config": {"workload": workload_name(args.frames, args.batch, args.new_tokens),
                       "sample": "each step is BASELINE configs[0] (2 frames, context ~580, 16 new tokens, bs=1) on the host CPUs — NOT the GPU "
                                 "arm's batch; see cpu_baseline.sample"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class Workload:
    """One (T frames, B examples per GPU, new tokens) workload on this rank's replica: synthetic frames (device + pinned
    host mirror), prompt ids, and the step function — generate_batch + the one collective (gather of the ids)."""

    def __init__(self, model, cfg, rank, world, dev, T, B, new):
        import ctypes as C

        import torch

        from teochat_b200 import lib as L
        from teochat_b200.weights import tensor_seed
        self.model, self.cfg, self.world, self.dev = model, cfg, world, dev
        self.T, self.B, self.new = T, B, new
        self.ids = [make_prompt_ids(cfg, T) for _ in range(B)]
        self.S0 = len(self.ids[0]) - T + T * cfg.tokens_per_image
        img = cfg.vision.image_size
        # synthetic frames, distinct per rank/sample, generated on the device then mirrored to pinned host memory
        frames_dev = torch.empty(B, T, img, img, 3, dtype=torch.uint8, device=dev)
        L.check(model.lib.teo_init_u8_hash(frames_dev.data_ptr(), frames_dev.numel(), C.c_uint64(tensor_seed(1234 + rank + 1000 * T, "frames")),
                                           torch.cuda.current_stream().cuda_stream))
        self.frames_host = frames_dev.cpu().pin_memory()
        self.dev_list = [frames_dev[b] for b in range(B)]
        self.host_list = [self.frames_host[b] for b in range(B)]
        self.n_total = B * world

    def step(self, frames):
        from teochat_b200 import dist as TD
        outs = self.model.generate_batch(self.ids, frames_u8=frames, max_new_tokens=self.new, time_phases=True)
        packed = TD.pack_tokens(outs, self.B, self.new, self.dev)
        gathered = TD.gather_tokens(packed, self.n_total)          # the one collective (includes the D2H of the ids)
        return outs, gathered

    def sync_all(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def timed(self, frames, k):
        """k steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks."""
        import torch
        import torch.distributed as dist
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.model.launch_count()
        phases, gen = [], 0
        e0.record()
        for _ in range(k):
            outs, _ = self.step(frames)
            phases.append(dict(self.model.last_timings))
            gen += sum(len(o) for o in outs)
        e1.record()
        self.sync_all()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            g = torch.tensor([gen], dtype=torch.int64, device=self.dev)
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            gen = int(g.item())
        return ms, gen, phases, self.model.launch_count() - l0

    def h2d_bytes(self):
        B, S0 = self.B, self.S0
        return int(self.frames_host.numel() + 4 * (3 * B * S0 + 3 * B + 1 + B * 64))

    def summary(self, ms_dev, gen_dev, ph_dev, ms_e2e, gen_e2e, k):
        """throughput figures of one measured workload (whole job: × world for the per-phase rates)."""
        w = self.world
        return {
            "workload": workload_name(self.T, self.B, self.new, self.S0), "global_batch": self.n_total,
            "value": gen_dev / (ms_dev / 1e3), "ms_per_step": ms_dev / k,
            "e2e": {"value": gen_e2e / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": self.h2d_bytes(),
                    "d2h_bytes_per_step": int(self.B * self.new * 4), "ms_per_step": ms_e2e / k},
            "vit_frames_per_s": sum(p["frames"] for p in ph_dev) / (sum(p["vit_ms"] for p in ph_dev) / 1e3) * w,
            "decode_tokens_per_s": sum(p["batch"] * p["decode_steps"] for p in ph_dev) / (sum(p["decode_ms"] for p in ph_dev) / 1e3) * w,
            "prefill_tokens_per_s": sum(p["prefill_tokens"] for p in ph_dev) / (sum(p["prefill_ms"] for p in ph_dev) / 1e3) * w,
            "phases_ms": {kk: sum(p[kk] for p in ph_dev) / k for kk in ("vit_ms", "prefill_ms", "decode_ms")},
        }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from teochat_b200 import dist as TD
    from teochat_b200.config import TeoConfig
    from teochat_b200.engine import TeoModel
    from teochat_b200.weights import TeoWeights

    rank, world, local = TD.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = TeoConfig.tiny() if args.tiny else TeoConfig.full()
    model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
    B, T, new = args.batch, args.frames, args.new_tokens
    wl = Workload(model, cfg, rank, world, dev, T, B, new)

    for _ in range(args.warmup):
        wl.step(wl.dev_list)
    if args.profile:                       # one device-resident step for ncu launch lists; prints no bench line
        wl.sync_all()
        torch.cuda.profiler.start()            # ncu --profile-from-start off: only this step is captured
        wl.step(wl.dev_list)
        wl.sync_all()
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"profile_only": True, "phases_ms": model.last_timings}), flush=True)
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, gen_dev, ph_dev, launches_eager = wl.timed(wl.dev_list, args.steps)
    ms_e2e, gen_e2e, ph_e2e, _ = wl.timed(wl.host_list, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    head = wl.summary(ms_dev, gen_dev, ph_dev, ms_e2e, gen_e2e, args.steps)

    # launches: eager calls are counted by the handle (teo_launch_count); a graph replay relaunches the kernels of one decode
    # step, counted on the eager step that preceded the capture (fallback: the upper bound 14 per layer + 6)
    per_step = model.decode_step_launches or (1 + cfg.llama.num_hidden_layers * 14 + 5)
    replays = sum(p.get("graph_replays", 0) for p in ph_dev)
    gpu_launches = int(launches_eager + replays * per_step)

    # per-rank phase times (N > 1: the step time is the MAX over ranks, so the slowest rank's phases explain it)
    per_rank = None
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, **head["phases_ms"]})

    roof = decode_attention_roofline(model, cfg, B, wl.S0 + new // 2, dev) if not args.tiny else None

    # the other BASELINE configs, measured briefly on the same replica (N = 1 only; the headline numbers above are final by now)
    others = None
    if world == 1 and not args.tiny and not args.no_other_configs:
        others = {}
        for ci, shape in CONFIGS.items():
            if shape == (T, B, new):
                continue
            del wl
            torch.cuda.empty_cache()
            wl = Workload(model, cfg, rank, world, dev, *shape)
            for _ in range(2):
                wl.step(wl.dev_list)
            r = wl.timed(wl.dev_list, 3)
            e = wl.timed(wl.host_list, 3)
            o = wl.summary(r[0], r[1], r[2], e[0], e[1], 3)
            o["steps"], o["warmup"] = 3, 2
            if ci == 4:
                o["note"] = "configs[4] is bs=16 over 8 GPUs = this shape on every rank; run `--config 4` under torchrun for the 8-GPU number"
            others[f"configs[{ci}]"] = o

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.tiny:
        cb = cpu_reference_leg(*CPU_SAMPLE, 2, 1)

    if rank == 0:
        pk = peaks()
        vit_flops = 155.3e9 + 10.74e9
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": head["workload"], "global_batch": head["global_batch"], "seq_len": None, "parallelism": f"dp{world}",
                       "l2": "working set (13.5 GB weights + KV pages) >> 126 MB L2; no flush needed", "weights": "random-init (hash) bf16"},
            "vit_frames_per_s": head["vit_frames_per_s"], "decode_tokens_per_s": head["decode_tokens_per_s"],
            "prefill_tokens_per_s": head["prefill_tokens_per_s"], "phases_ms": head["phases_ms"],
            "vit_tensor_frac_of_measured_peak": (head["vit_frames_per_s"] / world) * vit_flops / (pk["bf16_tflops_sustained"] * 1e12),
            "vit_tensor_frac_of_burst_peak": (head["vit_frames_per_s"] / world) * vit_flops / (pk["bf16_tflops"] * 1e12),
            "e2e": head["e2e"], "gpu_launches": gpu_launches, "clocks": clocks, "peaks": pk,
        }
        line["config"]["seq_len"] = int(head["workload"].split("context ")[1].split(",")[0]) + new
        if per_rank:
            line["per_rank_phases_ms"] = per_rank
        if others:
            line["other_configs"] = others
        if roof:
            line["roofline"] = roof
        if cb:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of decode_attn_mma_kernel from the one `ncu --set full` capture at
# exactly (bs=32, S=2258): profiles/r01_decode_attn_mma.txt (1.208262 GB + 8.388864 MB; whole 64-token pages)
NCU_TRAFFIC = {(32, 2258): 1208262000 + 8388864}


def decode_attention_roofline(model, cfg, B, S, dev, iters=20):
    """Dominant kernel of the decode loop: decode_attn_mma_kernel (+ its split merge), launched exactly as
    teo_llama_decode_step launches it (with the handle).  Algorithmic bytes per launch =
    Σ_seq 2(K,V)·heads·head_dim·2 B·S = 16 384·S per sequence per layer (524 288·S
    per token over 32 layers); timed alone with CUDA events on the launching stream, rotating over 8
    layer-sized KV pools (8 × 1.2 GB ≫ 126 MB L2, so every launch streams from HBM)."""
    import ctypes as C

    import torch

    from teochat_b200 import lib as L
    l, ps = cfg.llama, cfg.kv_page_size
    H, hd = l.num_attention_heads, l.head_dim
    pages_per = (S + ps - 1) // ps
    n_layers_resident = 8                      # rotate over 8 layers' pools so consecutive launches never hit L2
    pool = torch.empty(n_layers_resident, B * pages_per, 2, H, ps, hd, dtype=torch.bfloat16, device=dev)
    pool.view(-1)[: 1 << 20].normal_()
    bt = torch.arange(B * pages_per, dtype=torch.int32, device=dev).view(B, pages_per)
    q = torch.randn(B, 3 * H * hd, device=dev).to(torch.bfloat16)
    sl = torch.full((B,), S, dtype=torch.int32, device=dev)
    out = torch.empty(B, H * hd, dtype=torch.bfloat16, device=dev)
    wsb = model.lib.teo_decode_attention_workspace_bytes(B, H, hd, 32)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def launch(i):
        L.check(model.lib.teo_decode_attention_h(model._h, q.data_ptr(), 3 * H * hd, pool[i % n_layers_resident].data_ptr(), bt.data_ptr(),
                                                 pages_per, sl.data_ptr(), out.data_ptr(), B, H, hd, ps, S, hd ** -0.5, ws.data_ptr(),
                                                 ws.numel(), stream))
    for i in range(8):
        launch(i)
    torch.cuda.synchronize(dev)
    blocks = []
    for _ in range(3):                  # three timed blocks of `iters` launches; the median block is reported (one block hit by a
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)     # host hiccup must not set the figure)
        e0.record()
        for i in range(iters):
            launch(i)
        e1.record()
        torch.cuda.synchronize(dev)
        blocks.append(e0.elapsed_time(e1) / iters)
    ms = sorted(blocks)[1]
    alg_bytes = B * 2 * H * hd * 2 * S
    pk = peaks()
    ach = alg_bytes / (ms / 1e3) / 1e9
    del pool
    kernel = "decode_attn_mma_kernel" if (hd, ps) == (128, 64) and os.environ.get("TEO_DEC_ATTN") is None else "decode_attn_persist_kernel"
    return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": ach / pk["hbm_gbs"], "traffic": NCU_TRAFFIC.get((B, S)) if kernel == "decode_attn_mma_kernel" else None,
            "traffic_source": "static: dram__bytes_read+write of one `ncu --set full` capture at exactly this (bs, S), "
                              "profiles/r01_decode_attn_mma.txt — not measured in this run",
            "peak_source": pk["source"] + " (burst copy)",
            "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms * 1e3,
            "us_per_launch_blocks": [b * 1e3 for b in blocks],
            "how": f"bs={B}, S={S}, median of 3 blocks of {iters} launches over 8 rotating layer pools, CUDA events"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="teochat_b200", choices=["teochat_b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[i] as the headline workload")
    ap.add_argument("--batch", type=int, default=None, help="examples per GPU (default: the config's)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--new-tokens", type=int, default=None)
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short configs[1] / configs[4] measurements at N=1")
    ap.add_argument("--tiny", action="store_true", help="tiny config (plumbing check only; not a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="run warm-up + exactly one step and exit (for ncu; not a bench number)")
    args = ap.parse_args()
    T, B, new = CONFIGS[args.config]
    args.frames = args.frames if args.frames is not None else T
    args.batch = args.batch if args.batch is not None else B
    args.new_tokens = args.new_tokens if args.new_tokens is not None else new
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
