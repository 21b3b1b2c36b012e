"""A/B of programmatic-dependent-launch masks on the decode loop (dev tool; not a bench number)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from bench import make_prompt_ids
    from teochat_b200.config import TeoConfig
    from teochat_b200.engine import TeoModel
    from teochat_b200.weights import TeoWeights
    cfg = TeoConfig.full(); dev = "cuda:0"
    model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
    B, T, new = 32, 8, 64
    ids = [make_prompt_ids(cfg, T) for _ in range(B)]
    frames = [torch.randint(0, 256, (T, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(B)]
    for _ in range(3):
        model.generate_batch(ids, frames_u8=frames, max_new_tokens=new, time_phases=True)
    ts = []
    for _ in range(4):
        model.generate_batch(ids, frames_u8=frames, max_new_tokens=new, time_phases=True)
        ts.append(model.last_timings["decode_ms"] / model.last_timings["decode_steps"])
    print(json.dumps({"mask": os.environ.get("TEO_PDL_MASK"), "graph": os.environ.get("TEO_NO_GRAPH", "0") != "1", "ms_per_step": ts}))
else:
    for mask in ["0", "7", "1", "4", "5", "3"]:
        env = dict(os.environ, TEO_PDL_MASK=mask)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
