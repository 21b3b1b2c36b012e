"""GPU parity tests, op level: every CUDA kernel against a plain PyTorch fp32 restatement of the
same op on the same (bf16-valued) inputs.  Tolerances are stated per test; integer / byte /
index work is bit-exact."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from helpers import bf, gemm, rel_err, rnd, stream
from teochat_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------ synthetic init (bit-exact vs oracle)
def test_hash_init_matches_oracle(teo):
    from oracle import hashinit as H
    from teochat_b200.weights import hash_scale, tensor_seed
    lib, _ = teo
    for name, n, std, mean in [("model.layers.0.self_attn.q_proj.weight", 100003, 0.02, 0.0), ("x.norm", 4096, 0.02, 1.0)]:
        seed = tensor_seed(1234, name)
        assert seed == H.tensor_seed(1234, name)
        out = torch.empty(n, dtype=torch.bfloat16, device=DEV)
        L.check(lib.teo_init_normal_hash_bf16(out.data_ptr(), n, C.c_uint64(seed), hash_scale(std), mean, stream()))
        ref = H.hash_normal((n,), seed, std, mean).to(torch.bfloat16)
        assert torch.equal(out.cpu(), ref)
        out32 = torch.empty(n, dtype=torch.float32, device=DEV)
        L.check(lib.teo_init_normal_hash_f32(out32.data_ptr(), n, C.c_uint64(seed), hash_scale(std), mean, stream()))
        assert torch.equal(out32.cpu(), H.hash_normal((n,), seed, std, mean))
    u = torch.empty(5000, dtype=torch.uint8, device=DEV)
    L.check(lib.teo_init_u8_hash(u.data_ptr(), 5000, C.c_uint64(77), stream()))
    assert torch.equal(u.cpu(), H.hash_u8((5000,), 77))


# ------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    # (M, N, K) — normal tiles (BN 256/128/64), ragged edges, swap-AB + split-K (M <= 128)
    (128, 256, 64), (256, 512, 128), (300, 1024, 1024), (257 * 3, 3072, 1024), (1000, 4096, 640),
    (200, 128, 256), (130, 64, 128), (513, 328, 200),
    (1, 4096, 4096), (5, 512, 256), (32, 12288, 4096), (32, 4096, 11008), (64, 1024, 512), (100, 32000, 512), (17, 256, 64),
    # long K over many rows: the 512-row pair tiles (gemm_pair_kernel<2>), M not a multiple of 512 / 256 / 128
    (8200, 768, 4096), (8977, 512, 4096), (8200, 768, 2048),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(teo, M, N, K):
    A, W = bf(rnd(M, K, seed=1)), bf(rnd(N, K, scale=K ** -0.5, seed=2))
    out = gemm(teo, A, W)
    ref = A.float() @ W.float().t()
    assert rel_err(out, ref) < 1e-2          # one bf16 rounding of the output: ≤ 2^-8 relative per element
    out32 = gemm(teo, A, W, out_fp32=True)
    assert rel_err(out32, ref) < 2e-5        # fp32 accumulate, order differs from torch


@pytest.mark.parametrize("M,N,K", [(514, 1024, 1024), (257, 4096, 1024), (32, 4096, 4096), (3, 512, 256), (100, 768, 320)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogues(teo, M, N, K, act):
    A, W = bf(rnd(M, K, seed=3)), bf(rnd(N, K, scale=K ** -0.5, seed=4))
    bias, res = bf(rnd(N, seed=5)), bf(rnd(M, N, seed=6))
    y = A.float() @ W.float().t() + bias.float()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    out = gemm(teo, A, W, bias=bias, act=act)
    assert rel_err(out, y) < 1e-2
    out = gemm(teo, A, W, bias=bias, residual=res, act=act)
    assert rel_err(out, y + res.float()) < 1e-2
    # in-place residual (C aliases residual), the way the model uses it
    x = res.clone()
    gemm(teo, A, W, bias=bias, residual=x, act=act, C_out=x)
    assert rel_err(x, y + res.float()) < 1e-2


@pytest.mark.parametrize("M,I,K", [(300, 512, 256), (257 * 3, 11008, 512), (129, 192, 128), (1000, 64 * 5, 320)])
@pytest.mark.parametrize("blocked", [False, True], ids=["rowmajor", "blocked"])
def test_gemm_swiglu_pairs_epilogue(teo, M, I, K, blocked):
    """TEO_ACT_SWIGLU_PAIRS: gate/up rows interleaved in blocks of 32, C = silu(gate)·up with N/2 columns — equal, bit for
    bit, to the unfused chain (GEMM → bf16 [gate|up] → teo_swiglu) because the epilogue rounds gate and up to bf16 first."""
    lib, h = teo
    if blocked and ((2 * I) % 128 or K % 64):
        pytest.skip("blocked layout needs N % 128 == 0 and K % 64 == 0")
    A = bf(rnd(M, K, seed=11))
    Wg, Wu = bf(rnd(I, K, scale=K ** -0.5, seed=12)), bf(rnd(I, K, scale=K ** -0.5, seed=13))
    W_cat = torch.cat([Wg, Wu]).contiguous()
    W_int = W_cat.view(2, I // 32, 32, K).permute(1, 0, 2, 3).contiguous().view(2 * I, K)
    gu = gemm(teo, A, W_cat)                                   # unfused: bf16 [M, 2I]
    want = torch.empty(M, I, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_swiglu(gu.data_ptr(), want.data_ptr(), M, I, stream()))
    ref = torch.nn.functional.silu(gu[:, :I].float()) * gu[:, I:].float()
    assert rel_err(want, ref) < 1e-2
    out = torch.full((M, I), float("nan"), dtype=torch.bfloat16, device=DEV)
    Wd = W_int
    if blocked:
        Wd = torch.empty_like(W_int)
        L.check(lib.teo_weight_to_blocked(W_int.data_ptr(), Wd.data_ptr(), 2 * I, K, stream()))
        L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wd.data_ptr(), out.data_ptr(), I, M, 2 * I, K, None, None, 0, L.ACT_SWIGLU_PAIRS, 0,
                                           None, 0, stream()))
    else:
        L.check(lib.teo_gemm_bf16(h, A.data_ptr(), K, Wd.data_ptr(), K, out.data_ptr(), I, M, 2 * I, K, None, None, 0, L.ACT_SWIGLU_PAIRS, 0,
                                  None, 0, stream()))
    assert torch.equal(out, want)
    # small M takes the swap-AB schedule, which has no such epilogue: refused, not silently wrong
    rc = lib.teo_gemm_bf16(h, A.data_ptr(), K, W_int.data_ptr(), K, out.data_ptr(), I, 32, 2 * I, K, None, None, 0, L.ACT_SWIGLU_PAIRS, 0,
                           None, 0, stream())
    assert rc == -4 and b"tiled schedule" in lib.teo_last_error()


def test_gemm_512_row_pair_tiles_with_epilogues(teo):
    """gemm_pair_kernel<2> (M >= 8192, K >= 4096: two 256-row sub-tiles per pair tile, one shared W k-block) with the epilogues the
    prefill uses on it: residual (o / down; the residual tiles are prefetched a chunk ahead, the first one before the accumulators
    are complete), SwiGLU pairs (gate/up), bias + activation."""
    lib, h = teo
    M, N, K = 8300, 1024, 4096
    A, W = bf(rnd(M, K, seed=3)), bf(rnd(N, K, scale=K ** -0.5, seed=4))
    bias, res = bf(rnd(N, scale=0.1, seed=5)), bf(rnd(M, N, seed=6))
    ref = A.float() @ W.float().t()
    assert rel_err(gemm(teo, A, W, bias=bias, residual=res), ref + bias.float() + res.float()) < 1e-2
    x = ref + bias.float()
    assert rel_err(gemm(teo, A, W, bias=bias, act=1), x * torch.sigmoid(1.702 * x)) < 1e-2
    inplace = res.clone()                                       # residual aliasing the output, as the o / down projections run
    gemm(teo, A, W, residual=inplace, C_out=inplace)
    assert rel_err(inplace, ref + res.float()) < 1e-2
    # SwiGLU pairs: rows interleaved in blocks of 32 (gate 32 | up 32)
    I = N // 2
    g_, u_ = W[:I].float(), W[I:].float()
    Wi = torch.stack([W[:I].view(I // 32, 32, K), W[I:].view(I // 32, 32, K)], dim=1).reshape(N, K).contiguous()
    gg, uu = bf(A.float() @ g_.t()).float(), bf(A.float() @ u_.t()).float()
    want = torch.nn.functional.silu(gg) * uu
    out = torch.empty(M, I, dtype=torch.bfloat16, device=A.device)
    L.check(lib.teo_gemm_bf16(h, A.data_ptr(), K, Wi.data_ptr(), K, out.data_ptr(), I, M, N, K, None, None, 0, 3, 0, None, 0, stream()), "swiglu pairs")
    assert rel_err(out, want) < 2e-2


def test_gemm_strided_operands(teo):
    """q/k/v-style views: leading dimension larger than the logical width."""
    big = bf(rnd(300, 3 * 256, seed=7))
    A = big[:, 256:512]
    W = bf(rnd(512, 256, scale=1 / 16, seed=8))
    out = gemm(teo, A, W)
    assert rel_err(out, A.float() @ W.float().t()) < 1e-2


def test_gemm_bad_args(teo):
    lib, h = teo
    A, W = bf(rnd(8, 60)), bf(rnd(8, 60))
    out = torch.empty(8, 8, dtype=torch.bfloat16, device=DEV)
    rc = lib.teo_gemm_bf16(h, A.data_ptr(), 60, W.data_ptr(), 60, out.data_ptr(), 8, 8, 8, 60, None, None, 0, 0, 0, None, 0, stream())
    assert rc == -1 and b"multiples of 8" in lib.teo_last_error()


# ------------------------------------------------------------------ patchify / embeddings / norms
@pytest.mark.parametrize("image,patch", [(224, 14), (56, 14)])
def test_patchify(teo, image, patch):
    from oracle import model as OM
    lib, _ = teo
    n, g = 3, image // patch
    kpad = (3 * patch * patch + 63) // 64 * 64
    frames = torch.randint(0, 256, (n, image, image, 3), dtype=torch.uint8, device=DEV)
    out = torch.empty(n * g * g, kpad, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_patchify_u8_nhwc(frames.data_ptr(), out.data_ptr(), n, image, patch, kpad, stream()))
    px = OM.normalize_u8_nhwc(frames.cpu())                        # [n,3,H,W] f32
    ref = torch.nn.functional.unfold(px, kernel_size=patch, stride=patch).transpose(1, 2).reshape(n * g * g, -1)
    assert torch.equal(out[:, :3 * patch * patch].cpu(), ref.to(torch.bfloat16))     # bit-exact
    assert (out[:, 3 * patch * patch:] == 0).all()
    out2 = torch.empty_like(out)
    L.check(lib.teo_patchify_f32_nchw(px.to(DEV).contiguous().data_ptr(), out2.data_ptr(), n, image, patch, kpad, stream()))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows,d", [(7, 128), (257 * 2, 1024), (33, 4096), (5, 256)])
def test_layernorm_rmsnorm(teo, rows, d):
    lib, _ = teo
    x, w, b = bf(rnd(rows, d, scale=2.0, seed=1)), bf(1 + 0.1 * rnd(d, seed=2)), bf(0.1 * rnd(d, seed=3))
    y = torch.empty_like(x)
    L.check(lib.teo_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, d, 1e-5, stream()))
    ref = torch.nn.functional.layer_norm(x.float(), (d,), w.float(), b.float(), 1e-5)
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()      # one bf16 rounding
    L.check(lib.teo_rmsnorm(x.data_ptr(), w.data_ptr(), y.data_ptr(), rows, d, 1e-5, stream()))
    xf = x.float()
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


def test_vit_assemble_and_drop_cls(teo):
    lib, _ = teo
    n, npch, d = 3, 16, 128
    po, cls, pos = bf(rnd(n * npch, d, seed=1)), bf(rnd(d, seed=2)), bf(rnd(npch + 1, d, seed=3))
    w, b = bf(1 + 0.1 * rnd(d, seed=4)), bf(0.1 * rnd(d, seed=5))
    hid = torch.empty(n * (npch + 1), d, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_vit_assemble_preln(po.data_ptr(), cls.data_ptr(), pos.data_ptr(), w.data_ptr(), b.data_ptr(), hid.data_ptr(),
                                       n, npch, d, 1e-5, stream()))
    emb = torch.cat([cls.float().expand(n, 1, d), po.float().view(n, npch, d)], 1) + pos.float()[None]
    ref = torch.nn.functional.layer_norm(emb, (d,), w.float(), b.float(), 1e-5).view(-1, d)
    assert (hid.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    feats = torch.empty(n, npch, d, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_vit_drop_cls(hid.data_ptr(), feats.data_ptr(), n, npch, d, stream()))
    assert torch.equal(feats, hid.view(n, npch + 1, d)[:, 1:])


def test_swiglu_splice_argmax(teo):
    lib, _ = teo
    rows, inter = 37, 512
    gu = bf(rnd(rows, 2 * inter, scale=2.0, seed=1))
    out = torch.empty(rows, inter, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_swiglu(gu.data_ptr(), out.data_ptr(), rows, inter, stream()))
    ref = torch.nn.functional.silu(gu[:, :inter].float()) * gu[:, inter:].float()
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # splice gather: bit-exact copy
    d = 256
    E, F = bf(rnd(100, d, seed=2)), bf(rnd(40, d, seed=3))
    src = torch.tensor([1, 5, -1, -2, -40, 99, 0, -7], dtype=torch.int32, device=DEV)
    o = torch.empty(len(src), d, dtype=torch.bfloat16, device=DEV)
    L.check(lib.teo_splice_embed(E.data_ptr(), F.data_ptr(), src.data_ptr(), o.data_ptr(), len(src), d, stream()))
    ref = torch.stack([E[s] if s >= 0 else F[-(s + 1)] for s in src.tolist()])
    assert torch.equal(o, ref)
    # greedy argmax + stop rule: ties → lowest index; finished rows keep emitting eos / pad
    B, V, max_new = 4, 32000, 8
    lg = rnd(B, V, seed=4)
    lg[0, 777] = 50.0; lg[0, 31999] = 50.0          # tie → 777
    lg[1, 2] = 60.0                                  # eos
    lg[2, 31999] = 70.0
    fin = torch.tensor([0, 0, 0, 1], dtype=torch.uint8, device=DEV)
    toks = torch.full((B, max_new), -1, dtype=torch.int32, device=DEV)
    nxt = torch.empty(B, dtype=torch.int32, device=DEV)
    L.check(lib.teo_argmax_step(lg.data_ptr(), V, fin.data_ptr(), toks.data_ptr(), max_new, 3, nxt.data_ptr(), B, 2, stream()))
    assert toks[:, 3].tolist() == [777, 2, 31999, -1]
    assert nxt.tolist() == [777, 2, 31999, 2]
    assert fin.tolist() == [0, 1, 0, 1]
    assert torch.equal(torch.argmax(lg[2]), torch.tensor(31999, device=DEV))


# ------------------------------------------------------------------ RoPE + KV pages
def _rope_tables(max_pos, hd, theta=10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None]
    return fr.cos().contiguous().to(DEV), fr.sin().contiguous().to(DEV)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], -1)


@pytest.mark.parametrize("H,hd", [(4, 128), (3, 64), (2, 40)], ids=["vec128", "vec64", "scalar40"])
def test_rope_kv_write(teo, H, hd):
    """16-byte kernel (head_dim % 16 == 0) and the scalar one (other head dims) against the rotate-half formula."""
    lib, _ = teo
    ps = 16
    lens = [5, 37, 16]
    T, B = sum(lens), len(lens)
    max_pages = 4
    n_pages = B * max_pages
    qkv = bf(rnd(T, 3 * H * hd, seed=1))
    orig = qkv.clone()
    pos = torch.cat([torch.arange(n) for n in lens]).to(torch.int32).to(DEV)
    sid = torch.cat([torch.full((n,), i) for i, n in enumerate(lens)]).to(torch.int32).to(DEV)
    perm = torch.randperm(n_pages)[: B * max_pages].view(B, max_pages).to(torch.int32).to(DEV)   # scattered pages
    pages = torch.zeros(n_pages, 2, H, ps, hd, dtype=torch.bfloat16, device=DEV)
    cos, sin = _rope_tables(64, hd)
    L.check(lib.teo_rope_kv_write(qkv.data_ptr(), pos.data_ptr(), sid.data_ptr(), pages.data_ptr(), perm.data_ptr(), max_pages, T, H,
                                  hd, ps, cos.data_ptr(), sin.data_ptr(), stream()))
    q, k, v = [orig[:, i * H * hd:(i + 1) * H * hd].float().view(T, H, hd) for i in range(3)]
    c = torch.cat([cos, cos], -1)[pos.long()][:, None]
    s = torch.cat([sin, sin], -1)[pos.long()][:, None]
    q_ref, k_ref = bf(q * c + _rotate_half(q) * s), bf(k * c + _rotate_half(k) * s)
    got_q = qkv[:, :H * hd].view(T, H, hd)
    got_k = qkv[:, H * hd:2 * H * hd].view(T, H, hd)
    # fp32 products may be contracted to FMA on the device: allow one bf16 ulp on a few elements
    assert (got_q.float() - q_ref.float()).abs().max().item() <= 2 ** -7 * q_ref.float().abs().max().item()
    assert (got_k.float() - k_ref.float()).abs().max().item() <= 2 ** -7 * k_ref.float().abs().max().item()
    assert torch.equal(qkv[:, 2 * H * hd:], orig[:, 2 * H * hd:])           # v untouched
    for t in range(T):
        b, p = int(sid[t]), int(pos[t])
        page, slot = int(perm[b, p // ps]), p % ps
        assert torch.equal(pages[page, 0, :, slot], got_k[t])               # cache holds exactly what attention will read
        assert torch.equal(pages[page, 1, :, slot], orig[t, 2 * H * hd:].view(H, hd))


# ------------------------------------------------------------------ attention
def _attn_ref(q, k, v, scale, causal):
    """q,k,v [S,H,hd] f32 → [S,H,hd]; mirrors the build's rounding of P to bf16 before PV."""
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    if causal:
        S = q.shape[0]
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool, device=q.device).triu(1)[None], float("-inf"))
    m = s.amax(-1, keepdim=True)
    p = torch.exp(s - m)
    o = torch.einsum("hqk,khd->qhd", p.to(torch.bfloat16).float(), v) / p.sum(-1).transpose(0, 1)[..., None]
    return o


@pytest.mark.parametrize("hd,H,lens,causal", [
    (64, 16, [257, 257, 257], False),        # ViT block shape
    (64, 2, [17, 17], False),                # tiny ViT
    (128, 4, [580], True),                   # config (1) prefill shape
    (128, 2, [1, 63, 64, 65, 130, 300], True),   # ragged batch with edge lengths
    (128, 2, [200, 31], False),
])
def test_flash_attention(teo, hd, H, lens, causal):
    lib, _ = teo
    T = sum(lens)
    qkv = bf(rnd(T, 3 * H * hd, seed=11))
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    out = torch.empty(T, H * hd, dtype=torch.bfloat16, device=DEV)
    scale = hd ** -0.5
    d = H * hd
    L.check(lib.teo_flash_attention(qkv.data_ptr(), 3 * d, qkv[:, d:].data_ptr(), 3 * d, qkv[:, 2 * d:].data_ptr(), 3 * d,
                                    out.data_ptr(), d, cu.data_ptr(), len(lens), max(lens), H, hd, scale, int(causal), stream()))
    o = 0
    for n in lens:
        q, k, v = [qkv[o:o + n, i * d:(i + 1) * d].float().view(n, H, hd) for i in range(3)]
        ref = _attn_ref(q, k, v, scale, causal).reshape(n, d)
        err = (out[o:o + n].float() - ref).abs().max().item()
        assert err <= 1.5e-2 * ref.abs().max().item(), (n, err)     # bf16 P and bf16 output rounding
        o += n


@pytest.mark.parametrize("hd,H,lens,causal,q_offset", [
    (64, 16, [257, 257, 257], False, 1),         # ViT block shape: CLS row by the row kernel + two query tiles
    (64, 16, [257, 257], False, 0),              # same tokens, everything tiled (third tile holds one row)
    (64, 2, [17, 17], False, 1),                 # tiny ViT
    (64, 3, [1, 33, 128, 129, 300, 513], False, 0),
    (64, 2, [100, 256, 700], True, 0),
    (128, 4, [580], True, 0),                    # config (1) prefill shape
    (128, 2, [1, 63, 64, 65, 130, 300], True, 0),    # ragged batch with edge lengths
    (128, 2, [127, 128, 129, 255, 256, 257, 385], True, 0),   # tile / block boundaries
    (128, 32, [2151, 2130], True, 0),            # bench prefill shape (two sequences, all heads)
    (128, 2, [200, 31, 1000], False, 0),
    (128, 2, [200, 31, 1000, 1], False, 1),       # row-0 warp at head_dim 128, incl. a one-token sequence
])
def test_flash_attention_tc(teo, hd, H, lens, causal, q_offset):
    """tcgen05 flash attention (the path teo_vit_encode / teo_llama_prefill take) vs the fp32 restatement."""
    lib, h = teo
    T = sum(lens)
    d = H * hd
    qkv = bf(rnd(T, 3 * d, seed=13))
    qkv[:, :d] *= 3.0                           # wider score range: exercises the running-max / lazy-rescale path
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    out = torch.full((T, d), float("nan"), dtype=torch.bfloat16, device=DEV)
    scale = hd ** -0.5
    L.check(lib.teo_flash_attention_tc(h, qkv.data_ptr(), 3 * d, qkv[:, d:].data_ptr(), 3 * d, qkv[:, 2 * d:].data_ptr(), 3 * d,
                                       out.data_ptr(), d, cu.data_ptr(), len(lens), max(lens), T, H, hd, scale, int(causal), q_offset,
                                       stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any(), "rows left unwritten"
    o = 0
    for n in lens:
        q, k, v = [qkv[o:o + n, i * d:(i + 1) * d].float().view(n, H, hd) for i in range(3)]
        ref = _attn_ref(q, k, v, scale, causal).reshape(n, d)
        err = (out[o:o + n].float() - ref).abs().max().item()
        assert err <= 1.5e-2 * ref.abs().max().item(), (n, err)     # bf16 P and bf16 output rounding
        o += n


def test_flash_attention_tc_matches_mma_path(teo):
    """The two flash kernels agree to bf16 rounding on the ViT shape (same rounding points)."""
    lib, h = teo
    hd, H, lens = 64, 16, [257] * 8
    T, d = sum(lens), H * hd
    qkv = bf(rnd(T, 3 * d, seed=14))
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    a = torch.empty(T, d, dtype=torch.bfloat16, device=DEV)
    b = torch.empty_like(a)
    args = (qkv.data_ptr(), 3 * d, qkv[:, d:].data_ptr(), 3 * d, qkv[:, 2 * d:].data_ptr(), 3 * d)
    L.check(lib.teo_flash_attention(*args, a.data_ptr(), d, cu.data_ptr(), len(lens), 257, H, hd, hd ** -0.5, 0, stream()))
    L.check(lib.teo_flash_attention_tc(h, *args, b.data_ptr(), d, cu.data_ptr(), len(lens), 257, T, H, hd, hd ** -0.5, 0, 1, stream()))
    assert (a.float() - b.float()).abs().max().item() <= 2 ** -7 * a.float().abs().max().item()


@pytest.mark.parametrize("with_handle", [False, True], ids=["cuda_cores", "handle"])
@pytest.mark.parametrize("hd,ps,H,lens", [
    (128, 64, 32, [2151, 580, 64, 65, 1, 4000]),      # forces several KV splits
    (128, 64, 4, [300] * 40),                         # many sequences → single split
    (128, 64, 3, [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 193]),   # page / sub-block boundaries, odd page counts
    (128, 16, 2, [1, 15, 16, 17, 100]),               # tiny-config page size
])
def test_decode_attention(teo, hd, ps, H, lens, with_handle):
    """Paged decode attention: `teo_decode_attention` (CUDA-core kernel) and `teo_decode_attention_h` (with the handle:
    the mma.sync / TMA kernel for head_dim 128, page 64 — what the decode step runs).  Cache rows past each sequence's
    length are poisoned with NaN: they must never reach the result."""
    lib, h = teo
    B = len(lens)
    max_pages = max((n + ps - 1) // ps for n in lens)
    n_pages = B * max_pages
    bt = torch.randperm(n_pages).view(B, max_pages).to(torch.int32).to(DEV)
    pages = bf(rnd(n_pages, 2, H, ps, hd, seed=21))
    for b, n in enumerate(lens):                       # poison everything this sequence does not own
        for j in range(max_pages):
            lo = max(0, min(ps, n - j * ps))
            pages[bt[b, j].item(), :, :, lo:, :] = float("nan")
    q = bf(rnd(B, 3 * H * hd, seed=22))                 # q rows inside a fused qkv buffer (ldq = 3*H*hd)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    out = torch.empty(B, H * hd, dtype=torch.bfloat16, device=DEV)
    wsb = lib.teo_decode_attention_workspace_bytes(B, H, hd, 32)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    scale = hd ** -0.5
    args = (q.data_ptr(), 3 * H * hd, pages.data_ptr(), bt.data_ptr(), max_pages, sl.data_ptr(), out.data_ptr(),
            B, H, hd, ps, max(lens), scale, ws.data_ptr(), ws.numel(), stream())
    L.check(lib.teo_decode_attention_h(h, *args) if with_handle else lib.teo_decode_attention(*args))
    torch.cuda.synchronize()
    for b, n in enumerate(lens):
        idx = bt[b, : (n + ps - 1) // ps].long()
        K = pages[idx, 0].permute(0, 2, 1, 3).reshape(-1, H, hd)[:n].float()     # [n,H,hd]
        V = pages[idx, 1].permute(0, 2, 1, 3).reshape(-1, H, hd)[:n].float()
        qq = q[b, : H * hd].float().view(1, H, hd)
        s = torch.einsum("qhd,khd->hqk", qq, K) * scale
        p = torch.exp(s - s.amax(-1, keepdim=True))
        ref = (torch.einsum("hqk,khd->qhd", p.to(torch.bfloat16).float(), V) / p.sum(-1).transpose(0, 1)[..., None]).reshape(-1)
        err = (out[b].float() - ref).abs().max().item()
        assert err <= 1.5e-2 * ref.abs().max().item(), (b, n, err)


def test_sample_step_distribution(teo):
    """Temperature + top-k sampling: only top-k ids are drawn, frequencies follow softmax(z/T) restricted to the
    top-k set, draws are reproducible per (seed, step, sequence) and T→0 degenerates to arg-max."""
    lib, _ = teo
    V, B, max_new, T, K = 1000, 8192, 4, 0.7, 5
    base = rnd(V, seed=5)
    lg = base[None].repeat(B, 1).contiguous()
    fin = torch.zeros(B, dtype=torch.uint8, device=DEV)
    toks = torch.full((B, max_new), -1, dtype=torch.int32, device=DEV)
    nxt = torch.empty(B, dtype=torch.int32, device=DEV)
    L.check(lib.teo_sample_step(lg.data_ptr(), V, T, K, C.c_uint64(123), fin.data_ptr(), toks.data_ptr(), max_new, 1, nxt.data_ptr(), B, 2,
                                stream()))
    draws = toks[:, 1].long()
    topv, topi = (base / T).topk(K)
    assert set(draws.tolist()) <= set(topi.tolist())
    want = torch.softmax(topv, 0)
    freq = torch.stack([(draws == i).float().mean() for i in topi])
    assert (freq - want).abs().max().item() < 0.02                  # 8192 draws: 3 sigma ≈ 0.017
    toks2 = torch.full((B, max_new), -1, dtype=torch.int32, device=DEV)
    fin.zero_()
    L.check(lib.teo_sample_step(lg.data_ptr(), V, T, K, C.c_uint64(123), fin.data_ptr(), toks2.data_ptr(), max_new, 1, nxt.data_ptr(), B, 2,
                                stream()))
    assert torch.equal(toks[:, 1], toks2[:, 1])                     # same (seed, step, seq) → same draw
    fin.zero_()
    L.check(lib.teo_sample_step(lg.data_ptr(), V, T, K, C.c_uint64(124), fin.data_ptr(), toks2.data_ptr(), max_new, 1, nxt.data_ptr(), B, 2,
                                stream()))
    assert not torch.equal(toks[:, 1], toks2[:, 1])
    fin.zero_()
    L.check(lib.teo_sample_step(lg.data_ptr(), V, 1e-4, 50, C.c_uint64(1), fin.data_ptr(), toks2.data_ptr(), max_new, 2, nxt.data_ptr(), B, 2,
                                stream()))
    assert (toks2[:, 2] == int(base.argmax())).all()
    # ties at the k-th value are all kept (HF masks scores < kth)
    tie = torch.full((4, 64), -5.0, device=DEV)
    tie[:, 10] = 1.0; tie[:, 20] = 0.5; tie[:, 30] = 0.5
    fin4 = torch.zeros(4, dtype=torch.uint8, device=DEV)
    t4 = torch.full((4, 2), -1, dtype=torch.int32, device=DEV)
    n4 = torch.empty(4, dtype=torch.int32, device=DEV)
    seen = set()
    for sd in range(40):
        fin4.zero_()
        L.check(lib.teo_sample_step(tie.data_ptr(), 64, 1.0, 2, C.c_uint64(sd), fin4.data_ptr(), t4.data_ptr(), 2, 0, n4.data_ptr(), 4, 2, stream()))
        seen |= set(t4[:, 0].tolist())
    assert seen == {10, 20, 30}


@pytest.mark.parametrize("M,N,K", [(300, 1024, 1024), (257 * 3, 3072, 640), (32, 4096, 4096), (5, 512, 256), (1000, 128, 192), (64, 22016, 4096)])
def test_gemm_blocked_weights(teo, M, N, K):
    """Blocked weight layout [N/128][K/64][128][64]: relayout is a pure permutation, and the GEMM gives the same
    result as with the row-major weights (normal tiles, swap-AB and split-K schedules)."""
    lib, h = teo
    A, W = bf(rnd(M, K, seed=1)), bf(rnd(N, K, scale=K ** -0.5, seed=2))
    bias, res = bf(rnd(N, seed=5)), bf(rnd(M, N, seed=6))
    Wb = torch.empty_like(W)
    L.check(lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K, stream()))
    assert torch.equal(Wb.view(N // 128, K // 64, 128, 64), W.view(N // 128, 128, K // 64, 64).permute(0, 2, 1, 3))
    ref = gemm(teo, A, W, bias=bias, residual=res, act=1)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    wsb = lib.teo_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
    L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb.data_ptr(), out.data_ptr(), N, M, N, K, bias.data_ptr(), res.data_ptr(), N, 1, 0,
                                       ws.data_ptr(), ws.numel(), stream()))
    assert torch.equal(out, ref)            # same tiles, same accumulation order: bit-identical
    rc = lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K + 8, stream())
    assert rc == -1


@pytest.mark.parametrize("h,w,s", [
    (300, 400, 224), (400, 300, 224),       # landscape / portrait: one side resized + cropped
    (1024, 1024, 224),                      # xBD / S2Looking tile: 4.57x antialiased downscale
    (100, 150, 224), (64, 64, 224),         # upscale
    (225, 224, 224), (224, 224, 224), (224, 500, 224),   # short side already 224: crop + normalise only (exact)
    (513, 333, 224), (897, 641, 224),       # odd widths: rows start at every 16-byte misalignment
    (240, 20000, 224),                      # 60 kB rows (> 48 KiB of shared memory)
    (77, 130, 56),                          # tiny-config image size
])
def test_resize_crop_normalize(teo, h, w, s):
    """`teo_resize_crop_normalize_u8` (through TeoImageProcessor.preprocess_device) against the host processor, which is
    the reference's chain on torch's own bicubic (processing_image.py:15-25), and against the oracle restatement."""
    from oracle import preprocess as OP
    from teochat_b200.processor import TeoImageProcessor
    rng = np.random.default_rng(h * 31 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    proc = TeoImageProcessor(s)
    want = proc.preprocess(img)["pixel_values"][0]
    got = proc.preprocess_device(img, DEV)[0].cpu()
    assert got.shape == want.shape == (3, s, s)
    err = (got - want).abs().max().item()
    assert err <= 2e-5, err                                   # fp32 summation order / min(std)
    if min(h, w) == s:
        assert torch.equal(got, want)
    if h * w <= 1024 * 1024:
        ora = torch.from_numpy(OP.preprocess_u8_hwc(img, s, proc.image_mean, proc.image_std))
        assert (got - ora).abs().max().item() <= 2e-5


def test_resize_feeds_the_tower_and_bad_args(teo):
    lib, _ = teo
    from teochat_b200.processor import TeoImageProcessor
    proc = TeoImageProcessor(224)
    imgs = [np.random.default_rng(i).integers(0, 256, (200 + 50 * i, 320, 3), dtype=np.uint8) for i in range(3)]
    a = proc.preprocess_device(imgs, DEV)
    b = torch.cat([proc.preprocess(im)["pixel_values"] for im in imgs])
    assert a.shape == (3, 3, 224, 224) and (a.cpu() - b).abs().max().item() <= 2e-5
    import ctypes as C
    m = (C.c_float * 3)(0.5, 0.5, 0.5)
    src = torch.zeros(100, 100, 3, dtype=torch.uint8, device=DEV)
    dst = torch.empty(3, 224, 224, device=DEV)
    ws = torch.empty(lib.teo_resize_workspace_bytes(100, 100, 224), dtype=torch.uint8, device=DEV)
    ok = (src.data_ptr(), 100, 100, 224, 224, 0, 0, 224, m, m, dst.data_ptr(), ws.data_ptr(), ws.numel(), stream())
    assert lib.teo_resize_crop_normalize_u8(*ok) == 0
    bad_crop = ok[:5] + (1,) + ok[6:]
    assert lib.teo_resize_crop_normalize_u8(*bad_crop) < 0 and b"crop" in lib.teo_last_error()
    no_ws = ok[:11] + (None, 0, stream())
    assert lib.teo_resize_crop_normalize_u8(*no_ws) < 0 and b"workspace" in lib.teo_last_error()
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,act,blocked", [(514, 3072, 1024, 0, True), (300, 4096, 1024, 1, True), (257 * 3, 384, 128, 0, False),
                                               (1000, 128, 256, 2, False), (200, 64, 128, 0, False)])
def test_gemm_folded_layernorm_and_row_stats(teo, M, N, K, act, blocked):
    """teo_gemm_bf16_ex: LayerNorm folded into the consuming linear (HF CLIPEncoderLayer layer_norm1 → q/k/v, layer_norm2 → fc1,
    modeling_image.py:136-151) against act(LayerNorm(x)·Wᵀ + b) in fp32, with the row statistics coming (a) from teo_row_stats
    and (b) from the stats_out of a GEMM that produced x (residual epilogue) — the way teo_vit_encode chains them."""
    lib, h = teo
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) * 1.5 + 0.7).to(torch.bfloat16)            # non-zero mean: the folded form subtracts mean·Σw
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(torch.bfloat16)
    b = (torch.randn(N, generator=g) * 0.1).to(torch.bfloat16)
    gamma = (1.0 + 0.1 * torch.randn(K, generator=g)).to(torch.bfloat16)
    beta = (0.1 * torch.randn(K, generator=g)).to(torch.bfloat16)
    eps = 1e-5
    y = torch.nn.functional.layer_norm(x.float(), (K,), gamma.float(), beta.float(), eps)
    want = y @ w.float().T + b.float()
    want = want * torch.sigmoid(1.702 * want) if act == 1 else (torch.nn.functional.gelu(want) if act == 2 else want)
    # folded tensors exactly as TeoWeights.fold_vit_layernorm builds them
    wf = (w.float() * gamma.float()[None, :]).to(torch.bfloat16)
    c = wf.float().sum(1).contiguous().to(DEV)
    bf_ = (w.float() @ beta.float() + b.float()).contiguous().to(DEV)
    xd, wfd = x.to(DEV), wf.to(DEV)
    if blocked:
        wb = torch.empty_like(wfd)
        L.check(lib.teo_weight_to_blocked(wfd.data_ptr(), wb.data_ptr(), N, K, stream()))
        wfd = wb
    slots_in = 3
    stats = torch.full((M, slots_in, 2), float("nan"), device=DEV)
    L.check(lib.teo_row_stats(xd.data_ptr(), stats.data_ptr(), M, K, slots_in, stream()), "teo_row_stats")
    st = stats.cpu()
    assert torch.allclose(st[:, 0, 0], x.float().sum(1), rtol=1e-5, atol=1e-3) and torch.allclose(st[:, 0, 1], (x.float() ** 2).sum(1), rtol=1e-5)
    assert (st[:, 1:] == 0).all()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    o = L.GemmOpts(act=act, w_blocked=int(blocked), ln_stats=stats.data_ptr(), ln_c=c.data_ptr(), ln_bias=bf_.data_ptr(), ln_slots=slots_in, ln_eps=eps)
    L.check(lib.teo_gemm_bf16_ex(h, xd.data_ptr(), K, wfd.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(o), None, 0, stream()), "teo_gemm_bf16_ex")
    err = rel_err(out.cpu(), want)
    print(f"folded LayerNorm GEMM {M}x{N}x{K} act {act}: rel err vs fp32 LayerNorm+Linear {err:.2e}")
    assert err <= 2 ** -7
    # (b) statistics emitted by a producing GEMM: x2 = a·w2ᵀ + residual, then the folded GEMM on x2
    K2 = 192
    a = (torch.randn(M, K2, generator=g)).to(torch.bfloat16).to(DEV)
    w2 = (torch.randn(K, K2, generator=g) * K2 ** -0.5).to(torch.bfloat16).to(DEV)
    res = xd.clone()
    n_slots = lib.teo_gemm_stats_slots(M, K, K2)
    assert n_slots >= 2
    stats2 = torch.full((M, n_slots, 2), float("nan"), device=DEV)
    o2 = L.GemmOpts(residual=res.data_ptr(), ldr=K, stats_out=stats2.data_ptr())
    L.check(lib.teo_gemm_bf16_ex(h, a.data_ptr(), K2, w2.data_ptr(), K2, res.data_ptr(), K, M, K, K2, C.byref(o2), None, 0, stream()), "producer GEMM")
    x2 = res.float().cpu()                                              # the bf16 rows as stored
    s2 = stats2.cpu()
    assert torch.isfinite(s2).all()
    assert torch.allclose(s2[..., 0].sum(1), x2.sum(1), rtol=1e-5, atol=1e-2) and torch.allclose(s2[..., 1].sum(1), (x2 ** 2).sum(1), rtol=1e-4)
    o3 = L.GemmOpts(act=act, w_blocked=int(blocked), ln_stats=stats2.data_ptr(), ln_c=c.data_ptr(), ln_bias=bf_.data_ptr(), ln_slots=n_slots, ln_eps=eps)
    L.check(lib.teo_gemm_bf16_ex(h, res.data_ptr(), K, wfd.data_ptr(), K, out.data_ptr(), N, M, N, K, C.byref(o3), None, 0, stream()), "consumer GEMM")
    want2 = torch.nn.functional.layer_norm(x2, (K,), gamma.float(), beta.float(), eps) @ w.float().T + b.float()
    want2 = want2 * torch.sigmoid(1.702 * want2) if act == 1 else (torch.nn.functional.gelu(want2) if act == 2 else want2)
    err2 = rel_err(out.cpu(), want2)
    print(f"   … with statistics from the producing GEMM's epilogue ({n_slots} slots): rel err {err2:.2e}")
    assert err2 <= 2 ** -7
    # the folded path needs the tiled schedule: a small-M call is refused, not silently mis-served
    small = L.GemmOpts(ln_stats=stats.data_ptr(), ln_c=c.data_ptr(), ln_bias=bf_.data_ptr(), ln_slots=slots_in, ln_eps=eps)
    if N >= 256:
        ws = torch.empty(lib.teo_gemm_workspace_bytes(8, N, K) + 16, dtype=torch.uint8, device=DEV)
        assert lib.teo_gemm_bf16_ex(h, xd.data_ptr(), K, wfd.data_ptr(), K, out.data_ptr(), N, 8, N, K, C.byref(small), ws.data_ptr(), ws.numel(), stream()) != 0
