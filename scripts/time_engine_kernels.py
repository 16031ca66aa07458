"""Time the fused edge kernels on the bench's own engine state (GPU box): warm, back-to-back calls."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
import bench
from flashmd import _lib as L
from flashmd.engine import ForceField, SchNetWeights, prior_terms_from_system, random_schnet_tensors
from flashmd.neighbor_list import radius_graph_csr
class A: n_beads=269; batch=128; blocks=3
args=A()
dev=torch.device("cuda")
sysd,pos_np=bench.build_system(args,seed=0)
B,n=args.batch,args.n_beads
pos=torch.from_numpy(pos_np).reshape(B*n,3).to(dev).contiguous()
types=torch.from_numpy(sysd["atom_types"]).repeat(B).to(dev)
mol_ptr=(torch.arange(B+1)*n).to(dev)
w=SchNetWeights.from_flat(random_schnet_tensors(0,num_blocks=3),sysd["cutoff"],50,dev)
e0=radius_graph_csr(pos,mol_ptr,sysd["cutoff"],idx_dtype=torch.int32)["edge_index"].shape[1]
for exact in (True, False):
    ff=ForceField(w,[],types,mol_ptr,precision="w16a16",edge_capacity=int(1.35*e0)+4096, exact_cutoff_grad=exact)
    ff.compute(pos); torch.cuda.synchronize()
    ff._st=L.stream_ptr(); ff._n=0
    def t(fn,reps=20):
        for _ in range(3): fn()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
    print("exact",exact,"edges",ff.num_edges(),"cap",ff.cap,
          "fwd %.4f ms"%t(lambda: ff._filter_cfconv(1,ff.a[1],ff.m)),
          "bwd %.4f ms"%t(lambda: ff._filter_cfconv_bwd(1,ff.a[1],ff.g_m)),
          "bwd again %.4f ms"%t(lambda: ff._filter_cfconv_bwd(1,ff.a[1],torch.randn_like(ff.g_m)) if False else ff._filter_cfconv_bwd(1,ff.a[1],ff.g_m)))
    gm=ff.g_m; print("  |g_m| mean %.3e  |a| mean %.3e"%(gm.float().abs().mean().item(), ff.a[1].float().abs().mean().item()))
    gr=torch.randn(gm.shape,device=gm.device).to(gm.dtype)
    print("  bwd randn g_m %.4f ms"%t(lambda: ff._filter_cfconv_bwd(1,ff.a[1],gr)))
    # capacity == exact edge count (no slack tiles)
