"""Bit-identity check of the decode step with the gate/up reduction + SwiGLU inside the stream-K GEMM (TEO_SK_FUSE=1) against the
stand-alone glue kernel (TEO_SK_FUSE=0): run once per setting (the switch is read once per process), `python tools/sk_fuse_check.py dump <file>`,
then `python tools/sk_fuse_check.py cmp <a> <b>`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def dump(path):
    from oracle import weights as OW
    from test_gpu_model import _full_width, _model
    from teochat_b200.config import TeoConfig
    out = {}
    for case, cfg, B, n_new in (("tiny_b5", TeoConfig.tiny(), 5, 12), ("tiny_b40", TeoConfig.tiny(), 40, 12), ("full_d2_b3", _full_width(2, 1), 3, 8),
                                ("full_d2_b32", _full_width(2, 1), 32, 6)):
        model = _model(cfg, 99)
        model.set_decode_chain(False)
        ids = [[1, 17, -200, 5, 6, 30 + b % 7] + ([-200, 9] if b % 3 == 0 else []) for b in range(B)]
        frames = [OW.synthetic_frames_u8(2 if b % 3 == 0 else 1, cfg.vision.image_size, 200 + b % 11) for b in range(B)]
        ids_e, lg_e = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
        ids_g = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1)
        _, lg_e2 = model.generate_batch(ids, frames_u8=frames, max_new_tokens=n_new, eos_token_id=-1, return_logits=True)
        assert torch.equal(lg_e, lg_e2), f"{case}: not reproducible run to run"
        assert ids_e == ids_g, f"{case}: graph replay differs from eager"
        out[case] = (lg_e.cpu(), ids_e, model.decode_step_launches)
        print(case, "launches per step", model.decode_step_launches, "finite", bool(torch.isfinite(lg_e).all()))
        del model
        torch.cuda.empty_cache()
    torch.save(out, path)


def cmp(a, b):
    A, B = torch.load(a), torch.load(b)
    for k in A:
        same = torch.equal(A[k][0], B[k][0]) and A[k][1] == B[k][1]
        print(k, "BIT-IDENTICAL" if same else f"DIFFERENT max |diff| {(A[k][0] - B[k][0]).abs().max().item():.3e}", "launches", A[k][2], "vs", B[k][2])
        assert same


if __name__ == "__main__":
    dump(sys.argv[2]) if sys.argv[1] == "dump" else cmp(sys.argv[2], sys.argv[3])
