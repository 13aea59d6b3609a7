"""How much of the denoiser's error at the benchmark shape comes from WHICH tensor-memory accumulation - CPU experiment
(oracle + the accumulator model of tmem_truncation_model.py), test infrastructure.  One complex of 256 + 15 residues; the
selected Linear layers of the oracle are evaluated as the tcgen05 kernels evaluate them (3xFP16 split, small products first,
accumulator truncated after every MMA), everything else stays fp32; errors against the plain oracle.
    python tests/experiments/k1536_accumulation_effect.py    -> profiles/r2_k1536_accumulation_effect.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'experiments'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from oracle import pepflow_oracle as orc
import benchdata
from tmem_truncation_model import gemm_model
from make_golden_headline_inputs import headline_inputs
torch.set_num_threads(os.cpu_count())
sd=benchdata.reference_state_dict(114514)
batch=benchdata.synthetic_batch(1,256,15,seed=21)
enc=orc.encode(sd,batch)
inp=headline_inputs(enc,batch); inp['t']=inp['t'][:1]
keys=("t","rotmats_t","trans_t","angles_t","seqs_t","node_embed","edge_embed","generate_mask","res_mask")
ref=orc.ga_encoder_forward(sd,*[inp[k] for k in keys])
orig=orc.linear
def run(sel):
    def lin(x,w,b=None):
        K=x.shape[-1]
        if sel(K, w.shape[0]):
            y=torch.from_numpy(gemm_model(x.reshape(-1,K).numpy(), w.numpy(), 1)).reshape(*x.shape[:-1], w.shape[0])
            return y if b is None else y+b
        return orig(x,w,b)
    orc.linear=lin
    try: out=orc.ga_encoder_forward(sd,*[inp[k] for k in keys])
    finally: orc.linear=orig
    r=lambda a,b: float((a-b).abs().max()/b.abs().max())
    return r(out[0],ref[0]), r(out[1],ref[1]), r(out[3],ref[3])
lines = []
lines.append("only K=1536 (IPA linear_out) through the truncating-accumulator model: rot %.2e trans %.2e logits %.2e"%run(lambda K,N: K==1536))
lines.append("only the K=128, N=128 node layers through the model:                    rot %.2e trans %.2e logits %.2e"%run(lambda K,N: K==128 and N==128))
lines.append("all K=128 layers (incl. projections, heads) + K=1536 through the model: rot %.2e trans %.2e logits %.2e"%run(lambda K,N: K in (128,1536)))

print("\n".join(lines))
open(os.path.join(ROOT, "profiles", "r2_k1536_accumulation_effect.txt"), "w").write("# denoiser error (B=1, L=271) when only the named layers use the truncating tensor-memory accumulator model\n" + "\n".join(lines) + "\n")
