"""Where does the full-size bf16 error floor come from?  (test infrastructure: oracle code only, run on the CPU.)

BASELINE.json configs[0] (2 frames, context 576, random-init CLIP-L + LLaMA-2-7B, fixture config1_full.npz), step-0 logits of the
oracle under four rounding policies, against the fp32 oracle on 8 matmul threads:
  * fp32 oracle on 3 threads (accumulation order only)            -> is the network chaotic at fp32?
  * fp32 residual stream, bf16-rounded GEMM / attention operands  -> would an "fp32 residual" mode reach 1e-2?
  * every bf16 rounding point of the product path (policy "bf16")
Result recorded in profiles/r02_oracle_rounding_study.txt; it is why the exact mode splits GEMM operands into three bf16 terms
instead of merely keeping the residual in fp32.   python tests/golden/rounding_study.py   (~3 min, ~30 GB)"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as OM, weights as OW
from teochat_b200.config import TeoConfig
import torch.nn.functional as F, math
cfg = TeoConfig.full()
torch.set_num_threads(8)
t0=time.time(); sd = OW.make_state_dict(cfg, 1234); print('weights', time.time()-t0, flush=True)
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'config1_full.npz'))
ids = g['ids_0'].tolist(); nf, fs = g['frames_0'].tolist()
px = OM.normalize_u8_nhwc(OW.synthetic_frames_u8(nf, cfg.vision.image_size, fs))
feats32 = OM.encode_images(sd, cfg, px, 'fp32')
emb = OM.splice(sd, cfg, ids, feats32)
def run(policy, threads, resid_fp32=False):
    torch.set_num_threads(threads)
    lm = OM.LlamaOracle(sd, cfg, policy)
    if resid_fp32:
        # monkeypatch: residual adds unrounded -> emulate by custom forward
        r = lm.r
        l = lm.l
        def fwd(x):
            S,H,hd = x.shape[0], l.num_attention_heads, l.head_dim
            pos = torch.arange(0,S); cos,sin = OM._rope_tables(pos,hd,l.rope_theta)
            x = x.float()
            for i in range(l.num_hidden_layers):
                p=f"model.layers.{i}."
                y = lm._rms(x, sd[p+"input_layernorm.weight"])
                q = r(F.linear(y, sd[p+"self_attn.q_proj.weight"])).view(S,H,hd).transpose(0,1)
                k = r(F.linear(y, sd[p+"self_attn.k_proj.weight"])).view(S,H,hd).transpose(0,1)
                v = r(F.linear(y, sd[p+"self_attn.v_proj.weight"])).view(S,H,hd).transpose(0,1)
                q = r(q*cos[None]+OM._rotate_half(q)*sin[None]); k = r(k*cos[None]+OM._rotate_half(k)*sin[None])
                s = (q@k.transpose(-1,-2))/math.sqrt(hd)
                causal = torch.arange(S)[None,:] <= pos[:,None]
                s = s.masked_fill(~causal[None], float('-inf'))
                m = s.amax(-1,keepdim=True); pe = torch.exp(s-m)
                o = (r(pe)@v)/pe.sum(-1,keepdim=True)
                o = r(o.transpose(0,1).reshape(S,H*hd))
                x = x + F.linear(o, sd[p+"self_attn.o_proj.weight"])
                y = lm._rms(x, sd[p+"post_attention_layernorm.weight"])
                gg = r(F.linear(y, sd[p+"mlp.gate_proj.weight"])); u = r(F.linear(y, sd[p+"mlp.up_proj.weight"]))
                a = r(F.silu(gg)*u)
                x = x + F.linear(a, sd[p+"mlp.down_proj.weight"])
            x = x[-1:]
            var = x.pow(2).mean(-1,keepdim=True); y = x*torch.rsqrt(var+l.rms_norm_eps)*sd["model.norm.weight"]
            return F.linear(y, sd["lm_head.weight"])
        return fwd(emb)[0]
    return lm.forward(emb)[0]
def rel(a,b): return ((a-b).abs().max()/b.abs().max()).item()
t0=time.time(); a = run('fp32', 8); print('fp32/8', time.time()-t0, flush=True)
b = run('fp32', 3); print('fp32 8 vs 3 threads:', rel(b,a), flush=True)
c = run('bf16', 8, resid_fp32=True); print('bf16 GEMM inputs + fp32 residual vs fp32:', rel(c,a), flush=True)
d = run('bf16', 8); print('bf16 policy vs fp32:', rel(d,a), flush=True)
top2 = a.topk(2).values; print('fp32 margin rel', ((top2[0]-top2[1])/a.abs().max()).item(), 'argmax', int(a.argmax()), int(b.argmax()), int(c.argmax()), int(d.argmax()))
