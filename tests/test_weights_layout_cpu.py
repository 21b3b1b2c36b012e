"""Host-side weight re-layouts (teochat_b200/weights.py) on the CPU: the gate/up interleave that the SwiGLU GEMM epilogue
relies on (TEO_ACT_SWIGLU_PAIRS, include/teochat_b200.h) and the tile-blocked layout (teo_weight_to_blocked) are pure
permutations with the documented index maps."""
import torch

from oracle import weights as OW
from teochat_b200.config import TeoConfig
from teochat_b200.weights import TeoWeights


def _unblock(w):
    n, k = w.shape
    return w.view(n // 128, k // 64, 128, 64).permute(0, 2, 1, 3).reshape(n, k)


def test_gate_up_interleave_and_blocking_are_the_documented_permutations():
    cfg = TeoConfig.tiny()
    sd = OW.make_state_dict(cfg, 3)
    w = TeoWeights.from_state_dict(sd, cfg, "cpu")
    I, h = cfg.llama.intermediate_size, cfg.llama.hidden_size
    assert w.gate_up_interleaved and w.blocked["llama"]
    for i in range(cfg.llama.num_hidden_layers):
        gate = sd[f"model.layers.{i}.mlp.gate_proj.weight"].to(torch.bfloat16)
        up = sd[f"model.layers.{i}.mlp.up_proj.weight"].to(torch.bfloat16)
        rows = _unblock(w.t[f"llama.{i}.gate_up_w"])                      # interleaved row-major [2I, h]
        for r in (0, 31, 32, 63, 64, 100, 2 * I - 33, 2 * I - 1):         # row 64b+j: gate row 32b+j (j<32) | up row 32b+j-32
            b, j = divmod(r, 64)
            want = gate[32 * b + j] if j < 32 else up[32 * b + j - 32]
            assert torch.equal(rows[r], want), r
        back = rows.view(I // 32, 2, 32, h).permute(1, 0, 2, 3).reshape(2 * I, h)
        assert torch.equal(back, torch.cat([gate, up]))
        # the other fused matrices are only blocked
        q = sd[f"model.layers.{i}.self_attn.q_proj.weight"].to(torch.bfloat16)
        assert torch.equal(_unblock(w.t[f"llama.{i}.qkv_w"])[:h], q)
    assert w.interleave_gate_up() is w and w.to_blocked() is w             # idempotent


def test_interleave_is_skipped_when_the_width_does_not_allow_it():
    cfg = TeoConfig.tiny()
    cfg.llama.intermediate_size = 144                                      # 144 % 32 != 0 → [gate; up] stays, SwiGLU is a kernel
    sd = OW.make_state_dict(cfg, 3)
    w = TeoWeights.from_state_dict(sd, cfg, "cpu")
    assert not w.gate_up_interleaved and not w.blocked["llama"]            # 288 rows: not a multiple of 128 either
    gate = sd["model.layers.0.mlp.gate_proj.weight"].to(torch.bfloat16)
    assert torch.equal(w.t["llama.0.gate_up_w"][:144], gate)
