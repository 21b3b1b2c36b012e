"""Development check: a sample duplicated inside a batch must give bit-identical logits at every decode step, also across a
KV-page boundary (full-size model, T=16 frames, context 4202 -> 4228).  `python tools/dup_check.py` on a B200."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200.config import TeoConfig
from teochat_b200.engine import TeoModel
from teochat_b200.weights import TeoWeights
from oracle import weights as OW
import bench
cfg = TeoConfig.full(); dev = 'cuda:0'
model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
T = 16
ids = bench.make_prompt_ids(cfg, T)
fa, fb = OW.synthetic_frames_u8(T, 224, 41), OW.synthetic_frames_u8(T, 224, 42)
stacked = torch.cat([fa, fb, fa]).to(dev)
proj = model.encode_images(frames_u8=stacked)
print('projector dup equal:', torch.equal(proj[:T], proj[2*T:]), 'max diff', (proj[:T].float() - proj[2*T:].float()).abs().max().item())
outs, lg = model.generate_batch([ids, ids, ids], frames_u8=[fa, fb, fa], max_new_tokens=26, eos_token_id=-1, return_logits=True)
for s in range(20, lg.shape[1]):
    print('step', s, 'equal', torch.equal(lg[0, s], lg[2, s]), 'max diff', (lg[0, s] - lg[2, s]).abs().max().item())
print(outs[0] == outs[2])
