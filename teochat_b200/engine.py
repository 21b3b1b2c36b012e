"""Batched TEOChat inference engine over the C-ABI kernels.

Replaces, for the inference hot path, the reference's ``LlavaLlamaForCausalLM`` +
``LlavaMetaForCausalLM`` + HF ``generate`` stack (llava_llama.py:56-108, llava_arch.py:137-346,
SURVEY.md §3.2): per-frame CLIP-ViT encode → mlp2x_gelu projector → multimodal splice →
ragged batched LLaMA prefill into a paged KV cache → greedy decode with the whole step captured
in a CUDA graph (no per-token host synchronisation; the reference syncs every token in
KeywordsStoppingCriteria, mm_utils.py:94).

PyTorch is used for device memory, streams and CUDA-graph capture only; every FLOP runs in
teochat_b200/csrc through teochat_b200.lib (ctypes).  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import lib as L
from .config import TeoConfig
from .constants import IMAGE_TOKEN_INDEX
from .weights import TeoWeights


_NVTX_OK = True


def _nvtx(push: Optional[str]) -> None:
    """NVTX range per phase (vit / prefill / decode) for nsys / ncu timelines; a no-op cost without a profiler and switched off
    for good if the NVTX bindings are unavailable."""
    global _NVTX_OK
    if not _NVTX_OK:
        return
    try:
        if push is None:
            torch.cuda.nvtx.range_pop()
        else:
            torch.cuda.nvtx.range_push(push)
    except Exception:
        _NVTX_OK = False


def _cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


class TeoModel:
    """Duck-types what the reference's callers touch on ``model``: ``generate``,
    ``device``, ``config``, ``get_image_tower()``, settable ``model.video_tower``."""

    VIT_CHUNK_FRAMES = int(os.environ.get("TEO_VIT_CHUNK", "512"))     # frames per teo_vit_encode call (workspace ∝ this)

    def __init__(self, cfg: TeoConfig, weights: TeoWeights, device=None, precision: Optional[str] = None):
        """``precision``: "bf16" (default; bf16 activations / KV, the measured path) or "exact" — the parity mode of
        include/teochat_b200.h ("Exact mode"): fp32 activations, residual stream and KV pages, split-bf16 tensor-core GEMMs
        with fp32 accumulation; matches the fp32 oracle to ~1e-5 at full depth.  TEO_PRECISION overrides the default."""
        if not torch.cuda.is_available():
            raise L.TeoError("teochat_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.config = cfg                      # reference callers read model.config
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        precision = precision or os.environ.get("TEO_PRECISION", "bf16")
        if precision not in ("bf16", "exact"):
            raise ValueError(f"precision must be 'bf16' or 'exact', got {precision!r}")
        self.exact = precision == "exact"
        self.dtype = torch.float32 if self.exact else torch.bfloat16      # storage type of activations / KV pages
        self.w = weights
        self.lib = L.load()
        self.model = SimpleNamespace(video_tower=None)     # eval.py:31 sets model.model.video_tower = None
        h = C.c_void_p()
        L.check(self.lib.teo_create(self.device.index, C.byref(h)), "teo_create")
        self._h = h
        self._ws_cache: Dict[str, torch.Tensor] = {}
        self._kv_pool: Optional[torch.Tensor] = None
        self._decode_states: Dict[tuple, SimpleNamespace] = {}
        self._build_structs()
        self.use_graph = os.environ.get("TEO_NO_GRAPH", "0") != "1"
        self.retire_finished = os.environ.get("TEO_NO_RETIRE", "0") != "1"
        self.set_pdl(os.environ.get("TEO_NO_PDL", "0") != "1")
        self.last_timings: Dict[str, float] = {}
        self.decode_step_launches = 0          # counted on the eager step that precedes graph capture

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            self._release_kv_allocator()
            if getattr(self, "_h", None):
                self.lib.teo_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _release_kv_allocator(self):
        a = getattr(self, "_kv_alloc", None)
        if a:
            self.lib.teo_kv_destroy(a)
        self._kv_alloc = None

    def set_pdl(self, enabled: bool):
        """Programmatic dependent launch inside the decode step (default on; results are identical)."""
        self.use_pdl = bool(enabled)
        L.check(self.lib.teo_set_pdl(self._h, 1 if enabled else 0), "teo_set_pdl")

    def set_decode_chain(self, enabled: bool):
        """Persistent decode chain kernel vs one kernel per GEMM (the default); results are bit-identical.  Cached decode graphs
        were captured with the previous setting, so they are dropped."""
        L.check(self.lib.teo_set_decode_chain(self._h, 1 if enabled else 0), "teo_set_decode_chain")
        self._decode_states.clear()

    def get_image_tower(self):
        return self

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def launch_count(self) -> int:
        return int(self.lib.teo_launch_count(self._h))

    def _ws(self, name: str, nbytes: int) -> torch.Tensor:
        t = self._ws_cache.get(name)
        if t is None or t.numel() < nbytes:
            if t is not None:
                del self._ws_cache[name]
                del t
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws_cache[name] = t
        return t

    def _build_structs(self):
        cfg, t = self.cfg, self.w.t
        v, l = cfg.vision, cfg.llama
        p = lambda k: t[k].data_ptr()
        n_run = cfg.vit_layers_run
        self._vit_layers = (L.VitLayer * max(n_run, 1))()
        # folded LayerNorm (teo_gemm_bf16_ex) is built and tested but OFF by default: measured 46.4 vs 43.7 ms per 256 frames — the
        # K = 1024 GEMMs are epilogue-bound, and the folded epilogues cost more than the 46 LayerNorm passes they replace
        fold = (not self.exact) and getattr(self.w, "ln_folded", False) and os.environ.get("TEO_VIT_LN_FOLD", "0") == "1"
        for i in range(n_run):
            for f, _ in L.VitLayer._fields_:
                folded_field = f.endswith(("_wf", "_c", "_bf"))
                setattr(self._vit_layers[i], f, (p(f"vit.{i}.{f}") if fold else None) if folded_field else p(f"vit.{i}.{f}"))
        self._vit = L.VitModel(hidden=v.hidden_size, inter=v.intermediate_size, heads=v.num_attention_heads,
                               image=v.image_size, patch=v.patch_size, kpad=self.w.kpad,
                               act=L.ACT_BY_NAME[v.hidden_act], layers_run=n_run, eps=v.layer_norm_eps,
                               w_blocked=int(bool(self.w.blocked.get("vit"))), exact=int(self.exact),
                               patch_w=p("vit.patch_w"), cls=p("vit.cls"), pos=p("vit.pos"),
                               pre_ln_w=p("vit.pre_ln_w"), pre_ln_b=p("vit.pre_ln_b"), layers=self._vit_layers)
        if cfg.mm_projector_type != "mlp2x_gelu":
            raise ValueError(f"Unknown projector type: {cfg.mm_projector_type}")   # projector/builder.py:51
        self._proj = L.Projector(in_dim=v.hidden_size, hidden=l.hidden_size, w_blocked=int(bool(self.w.blocked.get("proj"))), exact=int(self.exact),
                                 w0=p("proj.w0"), b0=p("proj.b0"),
                                 w2=p("proj.w2"), b2=p("proj.b2"))
        # RoPE tables exactly as HF builds cos_cached/sin_cached (fp32, on the host)
        hd = l.head_dim
        self.rope_max_pos = max(l.max_position_embeddings, 8192)
        inv_freq = 1.0 / (l.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
        fr = torch.arange(self.rope_max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
        self._rope_cos = fr.cos().contiguous().to(self.device)
        self._rope_sin = fr.sin().contiguous().to(self.device)
        self._llama_layers = (L.LlamaLayer * l.num_hidden_layers)()
        for i in range(l.num_hidden_layers):
            for f in ("in_norm", "qkv_w", "o_w", "post_norm", "gate_up_w", "down_w"):
                setattr(self._llama_layers[i], f, p(f"llama.{i}.{f}"))
        self._llama = L.LlamaModel(hidden=l.hidden_size, inter=l.intermediate_size, heads=l.num_attention_heads,
                                   layers=l.num_hidden_layers, vocab=l.vocab_size, page_size=cfg.kv_page_size,
                                   rope_max_pos=self.rope_max_pos, eps=l.rms_norm_eps,
                                   w_blocked=int(bool(self.w.blocked.get("llama"))),
                                   gate_up_interleaved=int(bool(getattr(self.w, "gate_up_interleaved", False))), exact=int(self.exact),
                                   rope_cos=self._rope_cos.data_ptr(), rope_sin=self._rope_sin.data_ptr(),
                                   embed=p("llama.embed"), final_norm=p("llama.final_norm"),
                                   lm_head=p("llama.lm_head"), layer=self._llama_layers)

    def _ensure_kv(self, n_pages: int):
        l, ps = self.cfg.llama, self.cfg.kv_page_size
        if self._kv_pool is None or self._kv_pool.shape[1] < n_pages:
            self._kv_pool = None
            alloc = torch.zeros if os.environ.get("TEO_KV_ZERO") == "1" else torch.empty      # zeros: debugging aid only
            self._kv_pool = alloc(l.num_hidden_layers, n_pages, 2, l.num_attention_heads, ps, l.head_dim,
                                  dtype=self.dtype, device=self.device)
        for i in range(l.num_hidden_layers):
            self._llama_layers[i].kv_pages = self._kv_pool[i].data_ptr()

    # ------------------------------------------------------------------ vision
    def encode_images(self, frames_u8: Optional[torch.Tensor] = None, pixel_values: Optional[torch.Tensor] = None) -> torch.Tensor:
        """encode_images (llava_arch.py:137-140): tower (hidden_states[select_layer], CLS dropped)
        + projector.  frames_u8: u8 [n,H,W,3] on the device, or pixel_values: f32 [n,3,H,W]
        (already normalised).  Returns [n, tokens_per_image, llama_hidden] in the model's storage type (bf16; f32 in exact mode)."""
        cfg, v = self.cfg, self.cfg.vision
        if cfg.mm_vision_select_feature != "patch":
            raise ValueError(f"Unexpected select feature: {cfg.mm_vision_select_feature}")
        src = frames_u8 if frames_u8 is not None else pixel_values
        if src is None:
            raise ValueError("encode_images needs frames_u8 or pixel_values")
        if frames_u8 is not None:
            if frames_u8.dtype != torch.uint8 or tuple(frames_u8.shape[1:]) != (v.image_size, v.image_size, 3):
                raise ValueError(f"frames_u8 must be u8 [n,{v.image_size},{v.image_size},3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        else:
            if tuple(pixel_values.shape[1:]) != (3, v.image_size, v.image_size):
                raise ValueError(f"pixel_values must be [n,3,{v.image_size},{v.image_size}], got {tuple(pixel_values.shape)}")
            src = pixel_values.to(device=self.device, dtype=torch.float32)
        src = src.to(self.device).contiguous()
        n, npch, d, hl = src.shape[0], v.num_patches, v.hidden_size, cfg.llama.hidden_size
        out = torch.empty(n, npch, hl, dtype=self.dtype, device=self.device)
        esz = out.element_size()
        stream = self._stream()
        for s in range(0, n, self.VIT_CHUNK_FRAMES):
            c = min(self.VIT_CHUNK_FRAMES, n - s)
            wsb = self.lib.teo_vit_workspace_bytes(C.byref(self._vit), c)
            ws = self._ws("vit", wsb)
            feats = self._ws("vit_feats", c * npch * d * esz)
            chunk = src[s:s + c]
            L.check(self.lib.teo_vit_encode(self._h, C.byref(self._vit),
                                            chunk.data_ptr() if frames_u8 is not None else None,
                                            None if frames_u8 is not None else chunk.data_ptr(),
                                            c, feats.data_ptr(), ws.data_ptr(), ws.numel(), stream), "teo_vit_encode")
            pwb = self.lib.teo_projector_workspace_bytes(C.byref(self._proj), c * npch)
            pws = self._ws("proj", pwb)
            L.check(self.lib.teo_projector_mlp2x(self._h, C.byref(self._proj), feats.data_ptr(), c * npch,
                                                 out[s:s + c].data_ptr(), pws.data_ptr(), pws.numel(), stream),
                    "teo_projector_mlp2x")
        return out

    # ------------------------------------------------------------------ splice plan (host, integer)
    def plan_splice(self, input_ids: Sequence[Sequence[int]], images_per_sample: Sequence[int]):
        """Index plan of prepare_inputs_labels_for_multimodal (llava_arch.py:251-331) for a ragged
        batch: for every output row the source (token id ≥ 0, or -(feature_row+1) for image rows),
        with the reference's truncation to tokenizer_model_max_length (llava_arch.py:296-299)."""
        tpi = self.cfg.tokens_per_image
        maxlen = self.cfg.tokenizer_model_max_length
        srcs, lens = [], []
        img_base = 0
        for ids, n_img in zip(input_ids, images_per_sample):
            ids = np.asarray(list(ids), dtype=np.int64)
            is_img = ids == IMAGE_TOKEN_INDEX
            k = int(is_img.sum())
            if k > n_img:
                raise IndexError(f"sample has {k} <image> tokens but only {n_img} images")   # llava_arch.py:287 would IndexError
            bad = (ids < 0) & ~is_img
            if bad.any() or (ids >= self.cfg.llama.vocab_size).any():
                raise ValueError("token id outside the vocabulary")
            reps = np.where(is_img, tpi, 1)
            starts = np.cumsum(reps) - reps
            out = np.empty(int(reps.sum()), dtype=np.int64)
            out[starts[~is_img]] = ids[~is_img]
            img_ord = np.cumsum(is_img) - 1
            for pos in np.nonzero(is_img)[0]:
                r0 = (img_base + int(img_ord[pos])) * tpi
                out[starts[pos]:starts[pos] + tpi] = -(np.arange(r0, r0 + tpi) + 1)
            if maxlen is not None:
                out = out[:maxlen]
            srcs.append(out)
            lens.append(len(out))
            img_base += n_img
        return srcs, lens

    # ------------------------------------------------------------------ decode state / step
    RETIRE_FRACTION = 0.25          # compact the batch at a sync point once this share of its rows has finished

    def _decode_state(self, B: int, shape_key: tuple, dws: torch.Tensor) -> SimpleNamespace:
        """Device-resident state of the decode loop for one batch shape, with stable addresses so that the CUDA graph captured
        over it is reused across calls (and across the compactions of one call).  A cached entry is dropped when what its
        graph baked in has moved: the page pools (base pointer or pages per layer) or the workspace."""
        max_new_tokens, max_pages = shape_key[0], shape_key[1]
        l, dev = self.cfg.llama, self.device
        key = (B, *shape_key, self.use_pdl)
        pool_id = (self._kv_pool.data_ptr(), int(self._kv_pool.shape[1]))
        st = self._decode_states.get(key)
        if st is None or st.pool_id != pool_id or st.dws_ptr != dws.data_ptr():
            with torch.inference_mode(False):     # cached across calls: must stay ordinary tensors (callers may use inference_mode)
                st = SimpleNamespace(
                    B=B, d_len=torch.empty(B, dtype=torch.int32, device=dev), d_bt=torch.empty(B * max_pages, dtype=torch.int32, device=dev),
                    logits=torch.empty(B, l.vocab_size, dtype=torch.float32, device=dev),
                    finished=torch.empty(B, dtype=torch.uint8, device=dev), tokens=torch.empty(B, max_new_tokens, dtype=torch.int32, device=dev),
                    next_ids=torch.empty(B, dtype=torch.int32, device=dev), step_ptr=torch.empty(1, dtype=torch.int32, device=dev),
                    seed_dev=torch.empty(1, dtype=torch.int64, device=dev), graph=None, pool_id=pool_id, dws_ptr=dws.data_ptr())
            self._decode_states.pop(key, None)
            if len(self._decode_states) >= 8:
                self._decode_states.pop(next(iter(self._decode_states)))
            self._decode_states[key] = st
        return st

    def _decode_step(self, st: SimpleNamespace, shape_key: tuple, dws: torch.Tensor) -> None:
        max_new_tokens, max_pages, total_len, eos = shape_key[:4]
        L.check(self.lib.teo_llama_decode_step(self._h, C.byref(self._llama), st.next_ids.data_ptr(), st.d_len.data_ptr(),
                                               st.finished.data_ptr(), st.tokens.data_ptr(), max_new_tokens, st.step_ptr.data_ptr(), st.B,
                                               total_len, st.d_bt.data_ptr(), max_pages, st.logits.data_ptr(), eos, dws.data_ptr(),
                                               dws.numel(), self._stream()), "teo_llama_decode_step")

    # ------------------------------------------------------------------ batched generation
    @torch.no_grad()
    def generate_batch(self, input_ids: Sequence[Sequence[int]], frames_u8: Optional[Sequence[torch.Tensor]] = None,
                       pixel_values: Optional[Sequence[torch.Tensor]] = None, max_new_tokens: int = 256,
                       eos_token_id: Optional[int] = None, return_logits: bool = False, time_phases: bool = False,
                       temperature: float = 0.0, top_k: int = 50, seed: int = 0):
        """Greedy generation for a ragged batch of independent (image sequence, prompt) pairs.

        frames_u8[i]: u8 [T_i,H,W,3] (host or device) — or pixel_values[i]: f32 [T_i,3,H,W].
        Returns a list of per-sample new-token id lists (eos included if produced), like
        ``output_ids[0, input_ids.shape[1]:]`` in eval/inference.py:75.
        ``temperature > 0`` samples on the device (temperature + top-k + multinomial, the reference's default
        decode mode, inference.py:67-69; HF's default top_k=50); ``temperature <= 0`` is greedy.
        """
        cfg, l = self.cfg, self.cfg.llama
        B = len(input_ids)
        if B == 0:
            return []
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")
        eos = l.eos_token_id if eos_token_id is None else eos_token_id
        sampling = temperature is not None and temperature > 0
        L.check(self.lib.teo_set_sampling(self._h, float(temperature) if sampling else 0.0, int(top_k), C.c_uint64(seed & (2 ** 63 - 1))),
                "teo_set_sampling")
        imgs = frames_u8 if frames_u8 is not None else pixel_values
        if imgs is None or len(imgs) != B:
            raise ValueError("need one image stack per sample")
        dev, stream = self.device, self._stream()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if time_phases else None
        if ev:
            ev[0].record()
        # ---- vision: all frames of the batch in one go
        _nvtx("teo.vit+projector")
        per_sample = [int(x.shape[0]) for x in imgs]
        stacked = torch.cat([x.to(dev, non_blocking=True) for x in imgs], dim=0)
        proj = self.encode_images(frames_u8=stacked) if frames_u8 is not None else self.encode_images(pixel_values=stacked)
        if ev:
            ev[1].record()
        _nvtx(None)
        _nvtx("teo.prefill")
        # ---- splice plan (host integers) → device
        srcs, lens = self.plan_splice(input_ids, per_sample)
        T, max_len = int(sum(lens)), int(max(lens))
        total_len = max_len + max_new_tokens
        if total_len > self.rope_max_pos:
            raise ValueError(f"context {total_len} exceeds RoPE table ({self.rope_max_pos})")
        ps = cfg.kv_page_size
        # page plan + allocation through the C-ABI (teo_kv_plan / teo_kv_alloc, csrc/kv_pages.cu)
        lens_c, pages_c = (C.c_int * B)(*lens), (C.c_int * B)()
        mx_c, tot_c = C.c_int(), C.c_int()
        L.check(self.lib.teo_kv_plan(lens_c, B, max_new_tokens, ps, pages_c, C.byref(mx_c), C.byref(tot_c)), "teo_kv_plan")
        pages_per, max_pages, n_pages = list(pages_c), mx_c.value, tot_c.value
        cu = np.zeros(B + 1, dtype=np.int32)
        cu[1:] = np.cumsum(lens)
        meta = np.empty(3 * T + (B + 1) + B + B + B * max_pages, dtype=np.int32)
        o = 0
        v_src = meta[o:o + T]; o += T
        v_pos = meta[o:o + T]; o += T
        v_sid = meta[o:o + T]; o += T
        v_cu = meta[o:o + B + 1]; o += B + 1
        v_last = meta[o:o + B]; o += B
        v_len = meta[o:o + B]; o += B
        v_bt = meta[o:o + B * max_pages].reshape(B, max_pages); o += B * max_pages
        v_src[:] = np.concatenate(srcs)
        v_pos[:] = np.concatenate([np.arange(n, dtype=np.int32) for n in lens])
        v_sid[:] = np.repeat(np.arange(B, dtype=np.int32), lens)
        v_cu[:] = cu
        v_last[:] = cu[1:] - 1
        v_len[:] = lens
        v_bt[:] = 0
        self._release_kv_allocator()                      # (left behind only if a previous call raised half-way)
        alloc = C.c_void_p()
        L.check(self.lib.teo_kv_create(n_pages, C.byref(alloc)), "teo_kv_create")
        self._kv_alloc = alloc
        for b in range(B):
            got = self.lib.teo_kv_alloc(alloc, lens[b] + max_new_tokens, ps, v_bt[b].ctypes.data_as(C.POINTER(C.c_int)), max_pages)
            if got != pages_per[b]:
                L.check(got if got < 0 else -1, "teo_kv_alloc")
        self._ensure_kv(n_pages)
        meta_h = torch.from_numpy(meta).pin_memory()
        meta_d = meta_h.to(dev, non_blocking=True)
        d_src, d_pos, d_sid = meta_d[0:T], meta_d[T:2 * T], meta_d[2 * T:3 * T]
        o = 3 * T
        d_cu = meta_d[o:o + B + 1]; o += B + 1
        d_last = meta_d[o:o + B]; o += B
        d_len0 = meta_d[o:o + B]; o += B
        d_bt0 = meta_d[o:o + B * max_pages]
        h = l.hidden_size
        # ---- per-shape decode state with stable addresses, so the captured decode graph is reused across calls
        dws = self._ws("decode", self.lib.teo_llama_decode_workspace_bytes(C.byref(self._llama), B, total_len))
        shape_key = (max_new_tokens, max_pages, total_len, eos, sampling, float(temperature or 0.0) if sampling else 0.0,
                     int(top_k) if sampling else 0)
        st = self._decode_state(B, shape_key, dws)
        d_len, d_bt, logits = st.d_len, st.d_bt, st.logits
        finished, tokens, next_ids, step_ptr = st.finished, st.tokens, st.next_ids, st.step_ptr
        d_len.copy_(d_len0)
        d_bt.copy_(d_bt0)
        finished.zero_()
        tokens.fill_(-1)
        step_ptr.fill_(1)
        st.seed_dev.fill_(int(seed) & (2 ** 63 - 1))
        L.check(self.lib.teo_set_sampling_seed_device(self._h, st.seed_dev.data_ptr()), "teo_set_sampling_seed_device")
        x = self._ws("x", T * h * proj.element_size())
        splice = self.lib.teo_splice_embed_f32 if self.exact else self.lib.teo_splice_embed
        L.check(splice(self.w.t["llama.embed"].data_ptr(), proj.data_ptr(), d_src.data_ptr(), x.data_ptr(), T, h, stream), "teo_splice_embed")
        pwb = self.lib.teo_llama_prefill_workspace_bytes(C.byref(self._llama), T, B)
        pws = self._ws("prefill", pwb)
        L.check(self.lib.teo_llama_prefill(self._h, C.byref(self._llama), x.data_ptr(), T, d_cu.data_ptr(), d_pos.data_ptr(),
                                           d_sid.data_ptr(), d_last.data_ptr(), B, max_len, d_bt.data_ptr(), max_pages,
                                           logits.data_ptr(), pws.data_ptr(), pws.numel(), stream), "teo_llama_prefill")
        if sampling:
            L.check(self.lib.teo_sample_step(logits.data_ptr(), l.vocab_size, float(temperature), int(top_k), C.c_uint64(seed & (2 ** 63 - 1)),
                                             finished.data_ptr(), tokens.data_ptr(), max_new_tokens, 0, next_ids.data_ptr(), B, eos, stream),
                    "teo_sample_step")
        else:
            L.check(self.lib.teo_argmax_step(logits.data_ptr(), l.vocab_size, finished.data_ptr(), tokens.data_ptr(), max_new_tokens, 0,
                                             next_ids.data_ptr(), B, eos, stream), "teo_argmax_step")
        step_logits = [logits.clone()] if return_logits else None
        if ev:
            ev[2].record()

        _nvtx(None)
        _nvtx("teo.decode")
        # ---- decode: one captured step per batch shape, replayed; finished sequences are retired at the 32-step sync
        n_steps = max_new_tokens - 1
        done = 0
        use_graph = self.use_graph and not return_logits
        replays = 0
        rows = list(range(B))                                   # original example index of every current batch row
        result = np.full((B, max_new_tokens), -1, dtype=np.int32)
        retired_pages = 0
        while done < n_steps:
            if use_graph and st.graph is None and n_steps - done >= 5:
                l0 = self.launch_count()
                self._decode_step(st, shape_key, dws)           # first step eagerly (warms function attributes / tensor maps), then capture once
                self.decode_step_launches = self.launch_count() - l0       # kernels in one decode step (= one graph replay)
                done += 1
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                cap_stream = torch.cuda.Stream(device=dev)
                cap_stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.graph(graph, stream=cap_stream):
                    self._decode_step(st, shape_key, dws)
                st.graph = graph
                continue
            if use_graph and st.graph is not None:
                st.graph.replay()
                replays += 1
            else:
                self._decode_step(st, shape_key, dws)
                if return_logits:
                    step_logits.append(st.logits.clone())
            done += 1
            if done % 32 == 0 and done < n_steps:               # one D2H sync every 32 tokens
                fin = st.finished.cpu().numpy().astype(bool)
                if fin.all():
                    break
                if self.retire_finished and not return_logits and fin.sum() >= self.RETIRE_FRACTION * len(rows):
                    # Retire finished sequences (the reference stops each sample at its own </s>, inference.py:57-72): keep their
                    # tokens, give their pages back (teo_kv_free) and continue with a compacted batch — a smaller decode state
                    # (its own cached CUDA graph) whose rows are copies of the survivors' lengths, page tables and next ids.
                    toks_now = st.tokens.cpu().numpy()
                    bt_now = st.d_bt.view(len(rows), max_pages).cpu().numpy()
                    keep = np.nonzero(~fin)[0]
                    for r in np.nonzero(fin)[0]:
                        result[rows[r]] = toks_now[r]
                        npg = pages_per[rows[r]]
                        pg = np.ascontiguousarray(bt_now[r, :npg], dtype=np.int32)
                        L.check(self.lib.teo_kv_free(alloc, pg.ctypes.data_as(C.POINTER(C.c_int)), npg), "teo_kv_free")
                        retired_pages += npg
                    idx = torch.from_numpy(keep).to(dev)
                    dws = self._ws("decode", self.lib.teo_llama_decode_workspace_bytes(C.byref(self._llama), len(keep), total_len))
                    st2 = self._decode_state(len(keep), shape_key, dws)
                    st2.d_len.copy_(st.d_len[idx])
                    st2.d_bt.copy_(st.d_bt.view(len(rows), max_pages)[idx].reshape(-1))
                    st2.next_ids.copy_(st.next_ids[idx])
                    st2.tokens.copy_(st.tokens[idx])
                    st2.finished.zero_()
                    st2.step_ptr.copy_(st.step_ptr)
                    st2.seed_dev.copy_(st.seed_dev)
                    L.check(self.lib.teo_set_sampling_seed_device(self._h, st2.seed_dev.data_ptr()), "teo_set_sampling_seed_device")
                    rows = [rows[r] for r in keep]
                    st = st2
        if ev:
            ev[3].record()
        toks = st.tokens.cpu().numpy()
        for r, orig in enumerate(rows):
            result[orig] = toks[r]
        self._release_kv_allocator()
        _nvtx(None)
        if ev:
            torch.cuda.synchronize(dev)
            self.last_timings = {"vit_ms": ev[0].elapsed_time(ev[1]), "prefill_ms": ev[1].elapsed_time(ev[2]),
                                 "decode_ms": ev[2].elapsed_time(ev[3]), "frames": int(sum(per_sample)),
                                 "prefill_tokens": T, "decode_steps": done, "batch": B, "graph_replays": replays,
                                 "final_batch": len(rows), "retired_pages": retired_pages}
        outs: List[List[int]] = []
        for b in range(B):
            row = result[b]
            cut = len(row)
            neg = np.nonzero(row < 0)[0]
            if len(neg):
                cut = int(neg[0])
            outs.append([int(t) for t in row[:cut]])
        if return_logits:
            return outs, torch.stack(step_logits, dim=1)     # [B, steps, vocab]
        return outs

    # ------------------------------------------------------------------ HF-like single-sample API
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor = None, images=None, do_sample: bool = False, temperature: float = 0.0,
                 max_new_tokens: int = 256, use_cache: bool = True, stopping_criteria=None, **kwargs) -> torch.Tensor:
        """``model.generate`` as called at eval/inference.py:64-72: ``input_ids`` [1,L] with -200 at
        image slots, ``images`` a list of T tensors [3,H,W].  Returns [1, L+n] (prompt ids unexpanded,
        SURVEY.md §8 quirk 6)."""
        if input_ids is None or input_ids.dim() != 2 or input_ids.shape[0] != 1:
            raise ValueError("generate expects input_ids of shape [1, L]; use generate_batch for batches")
        sample_t = float(temperature) if (do_sample and temperature and temperature > 0) else 0.0
        eos = self.cfg.llama.eos_token_id
        # The eval path's ["</s>"] rule runs on the device (last id == eos).  Any other criterion (multi-token stop strings of
        # the llava_llama_2 / plain templates, user criteria) is evaluated on the host over the generated prefix afterwards —
        # same result as HF's per-token check, because step n of the decode does not depend on later steps.
        host_criteria = [sc for sc in (stopping_criteria or []) if not (hasattr(sc, "is_eos_only") and sc.is_eos_only(eos))]
        if isinstance(images, (list, tuple)):
            px = torch.stack([im.to(torch.float32) for im in images]) if len(images) else None
        else:
            px = images
        if px is None:
            raise ValueError("generate needs images (the text-only branch is not part of the TEOChat path)")
        if px.dim() == 3:
            px = px[None]
        ids = input_ids[0].tolist()
        out = self.generate_batch([ids], pixel_values=[px], max_new_tokens=max_new_tokens, temperature=sample_t,
                                  top_k=int(kwargs.get("top_k", 50) or 0), seed=int(kwargs.get("seed", 0)))[0]
        new = torch.tensor(out, dtype=input_ids.dtype, device=input_ids.device)[None]
        full = torch.cat([input_ids, new], dim=1)
        if host_criteria:
            L0 = input_ids.shape[1]
            for n in range(1, len(out) + 1):
                if any(bool(sc(full[:, :L0 + n], None)) for sc in host_criteria):       # HF stops when ANY criterion fires
                    return full[:, :L0 + n]
        return full
