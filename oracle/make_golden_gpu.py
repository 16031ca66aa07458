"""Golden vectors from the UNMODIFIED reference's own GPU (Triton) paths.  TEST INFRASTRUCTURE ONLY.

Run ON THE GPU BOX (the reference's Triton kernels need a GPU; /root/reference does not exist there, the
`pip install --target baseline/_ref` copy of the reference travels with the snapshot):

    python oracle/make_golden_gpu.py [--out gpurun_out/golden_gpu]      # then copy *.npz into tests/golden/

What runs: the reference package from baseline/_ref, imported behind oracle/shims (nvtx, torch_geometric,
torch_cluster, jsonargparse, ruamel.yaml), default flags — every MLCG_* toggle "1" (models/schnet.py:52-56) —
through `LangevinSimulation.attach_model_and_configurations` (simulation/base.py:319-368), which applies
`apply_gptq_w16a16_to_model` (models/gptq.py:373-460) for gptq="w16a16", then ONE model forward on the attached
initial data (simulation/base.py:864).  Nothing of this repository's library is on that path: `radius_graph` is the
pure-torch stand-in of oracle/shims/torch_cluster (same pairs/order as the committed CPU golden edge lists), `math` is
injected into flashmd.kernels.cfconv_kernels (the shipped fused-RBF backward raises NameError without it).

Inputs are the committed CPU golden files (weights + system arrays; tests/golden/schnet_*.npz, written by
oracle/make_golden.py from the reference's CPU path), plus one cfg2-shaped system (269 beads) with the n54 weights.
Outputs per case `w16a16_triton_<case>.npz`:
  w16.energy.* / w16.forces.*     reference W16A16 path (gptq="w16a16": fp16 filter + output networks, Triton CSR CFConv)
  tf32.energy.* / tf32.forces.*   reference "fp32" GPU path (gptq=None; tl.dot is TF32), for information
  edge_index                      the SchNet neighbour list of that forward
  pos / atom_types / cutoff       inputs (for the 269-bead case, which has no CPU golden file)
"""
import argparse
import importlib.util
import math
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, os.path.join(ROOT, "baseline", "_ref"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden_gpu"))
args = ap.parse_args()
os.makedirs(args.out, exist_ok=True)
assert torch.cuda.is_available(), "the reference's Triton path needs a GPU"

spec = importlib.util.spec_from_file_location(
    "fmd_synthetic", os.path.join(ROOT, "flash-molecular-dynamics_b200", "flashmd", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

import flashmd  # noqa: E402  (the reference)
assert "baseline/_ref" in flashmd.__file__, flashmd.__file__
import flashmd.kernels.cfconv_kernels as _ck  # noqa: E402
_ck.math = math
from flashmd.data import AtomicData  # noqa: E402
from flashmd.models import CosineCutoff, GaussianBasis, GradientsOut, StandardSchNet, SumOut  # noqa: E402
from flashmd.neighbor_list.neighbor_list import make_neighbor_list  # noqa: E402
from flashmd.prior import Dihedral, HarmonicAngles, HarmonicBonds, Repulsion  # noqa: E402
from flashmd.simulation import LangevinSimulation  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def statistics(g):
    ty = g["sys.atom_types"]
    sb, sa, sd, sr = {}, {}, {}, {}
    for i, j in g["sys.bonds"].T:
        key = (int(ty[i]), int(ty[j]))
        sb[key] = {"k": float(g["stats.bonds.k"][key]), "x_0": float(g["stats.bonds.x_0"][key])}
    for i, j, k in g["sys.angles"].T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]))
        sa[key] = {"k": float(g["stats.angles.k"][key]), "x_0": float(g["stats.angles.x_0"][key])}
    k1, k2, v0 = g["stats.dihedrals.k1_central"], g["stats.dihedrals.k2_central"], g["stats.dihedrals.v0_central"]
    nd = k1.shape[0]
    for i, j, k, l in g["sys.dihedrals"].T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]), int(ty[l]))
        c = (int(ty[j]), int(ty[k]))
        sd[key] = {"k1s": {f"k1_{q + 1}": float(k1[(q,) + c]) for q in range(nd)},
                   "k2s": {f"k2_{q + 1}": float(k2[(q,) + c]) for q in range(nd)}, "v_0": float(v0[c])}
    for i, j in g["sys.nonbonded"].T:
        key = (int(ty[i]), int(ty[j]))
        sr[key] = {"sigma": float(g["stats.repulsion.sigma"][key])}
    return sb, sa, sd, sr, nd


def build(g, with_priors=True):
    """Reference model objects with the weights of a CPU golden file."""
    hp = [int(v) for v in g["meta.hparams"]]
    hidden, filters, num_rbf, nblocks, widths = hp[0], hp[1], hp[2], hp[3], hp[4:]
    rc = float(g["sys.cutoff"])
    schnet = StandardSchNet(GaussianBasis(CosineCutoff(0.0, rc), num_rbf=num_rbf), CosineCutoff(0.0, rc),
                            output_hidden_layer_widths=list(widths), hidden_channels=hidden,
                            embedding_size=g["w.embedding"].shape[0], num_filters=filters, num_interactions=nblocks)
    W = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")}
    with torch.no_grad():
        schnet.embedding_layer.weight.copy_(W["embedding"])
        for l, blk in enumerate(schnet.interaction_blocks):
            cf = blk.conv
            cf.lin1.weight.copy_(W[f"b{l}.lin1_w"])
            cf.filter_network.layers[0].weight.copy_(W[f"b{l}.f0_w"])
            cf.filter_network.layers[0].bias.copy_(W[f"b{l}.f0_b"])
            cf.filter_network.layers[2].weight.copy_(W[f"b{l}.f1_w"])
            cf.lin2.weight.copy_(W[f"b{l}.lin2_w"]); cf.lin2.bias.copy_(W[f"b{l}.lin2_b"])
            blk.lin.weight.copy_(W[f"b{l}.lin_w"]); blk.lin.bias.copy_(W[f"b{l}.lin_b"])
        lin = [m for m in schnet.output_network.layers if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lin):
            m.weight.copy_(W[f"out{i}_w"])
            if m.bias is not None:
                m.bias.copy_(W[f"out{i}_b"])
    models = {"SchNet": GradientsOut(schnet)}
    if with_priors:
        sb, sa, sd, sr, nd = statistics(g)
        models |= {"bonds": GradientsOut(HarmonicBonds(sb)), "angles": GradientsOut(HarmonicAngles(sa)),
                   "dihedrals": GradientsOut(Dihedral(sd, n_degs=nd)), "repulsion": GradientsOut(Repulsion(sr))}
    model = SumOut(torch.nn.ModuleDict(models))
    configs = []
    for b in range(g["sys.pos"].shape[0]):
        nls = {}
        if with_priors:
            nls = {"bonds": make_neighbor_list("bonds", 2, torch.from_numpy(g["sys.bonds"])),
                   "angles": make_neighbor_list("angles", 3, torch.from_numpy(g["sys.angles"])),
                   "dihedrals": make_neighbor_list("dihedrals", 4, torch.from_numpy(g["sys.dihedrals"])),
                   "repulsion": make_neighbor_list("repulsion", 2, torch.from_numpy(g["sys.nonbonded"]))}
        configs.append(AtomicData.from_points(pos=torch.from_numpy(g["sys.pos"][b].copy()),
                                              atom_types=torch.from_numpy(g["sys.atom_types"]),
                                              masses=torch.from_numpy(g["sys.masses"]), neighborlist=nls))
    return model, schnet, configs


def evaluate(g, gptq, with_priors=True):
    model, schnet, configs = build(g, with_priors)
    tmp = tempfile.mkdtemp()
    sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=2, save_interval=2, export_interval=2,
                             random_seed=1, device="cuda", dtype="single", filename="g", output_dir=tmp,
                             specialize_priors=True, compile_model=False, gptq=gptq)
    sim.attach_model_and_configurations(model, configs, beta=1.67)
    data = sim.initial_data
    data.out = {}
    data._dump_neighbor_list = True          # models/schnet.py:365-367: keep the edge list of this forward
    data = sim.model(data)
    out = {}
    for name in sim.model.models.keys():
        out[f"energy.{name}"] = data.out[name]["energy"].detach().float().cpu().numpy().copy()
        out[f"forces.{name}"] = data.out[name]["forces"].detach().float().cpu().numpy().copy()
    out["energy.total"] = data.out["energy"].detach().float().cpu().numpy().copy()
    out["forces.total"] = data.out["forces"].detach().float().cpu().numpy().copy()
    out["edge_index"] = data.out["SchNet"]["edge_index"].cpu().numpy().copy()
    return out


def run_case(tag, g, extra=None):
    arrs = dict(extra or {})
    for mode, gptq in (("w16", "w16a16"), ("tf32", None)):
        try:
            out = evaluate(g, gptq)
        except Exception as ex:       # a case the reference itself cannot run (e.g. 2-layer output net under GPTQ)
            import traceback
            traceback.print_exc()
            arrs[f"{mode}.error"] = np.array(repr(ex)[:400])
            continue
        ei = out.pop("edge_index")
        if "edge_index" in arrs:
            assert np.array_equal(arrs["edge_index"], ei)
        arrs["edge_index"] = ei
        arrs |= {f"{mode}.{k}": v for k, v in out.items()}
    arrs["meta.torch"] = np.array(torch.__version__)
    import triton
    arrs["meta.triton"] = np.array(triton.__version__)
    arrs["meta.gpu"] = np.array(torch.cuda.get_device_name(0))
    np.savez_compressed(os.path.join(args.out, f"w16a16_triton_{tag}.npz"), **arrs)
    msg = {k: (v.shape if v.ndim else str(v)) for k, v in arrs.items() if "energy.total" in k or "error" in k}
    print(tag, "E =", arrs["edge_index"].shape[1] if "edge_index" in arrs else None, msg, flush=True)
    if "w16.forces.SchNet" in arrs and "ref64.forces.SchNet" in g:
        a, b = arrs["w16.forces.SchNet"], g["ref64.forces.SchNet"]
        print("   |w16 - cpu fp64 (exact cut-off gradient)| / |.| =", float(np.linalg.norm(a - b) / np.linalg.norm(b)))
        a = arrs.get("tf32.forces.SchNet")
        if a is not None:
            print("   |tf32 - cpu fp64| / |.| =", float(np.linalg.norm(a - b) / np.linalg.norm(b)))


if __name__ == "__main__":
    torch.set_float32_matmul_precision("high")     # scripts/nvt_langevin.py:38
    for name in ("schnet_n54_b4", "schnet_n24_b3_l2", "schnet_n40_b2_l5"):
        path = os.path.join(GOLDEN, name + ".npz")
        if not os.path.exists(path):
            print("skip", name)
            continue
        run_case(name, dict(np.load(path, allow_pickle=False)))
    # cfg2-shaped: 269 beads (the benchmark molecule), 2 molecules, weights of the n54 golden (same hyper-parameters)
    g = dict(np.load(os.path.join(GOLDEN, "schnet_n54_b4.npz"), allow_pickle=False))
    system = syn.synthetic_system(2, 269, seed=0)
    g2 = {k: v for k, v in g.items() if k.startswith(("w.", "meta.", "stats."))}
    for k in ("pos", "atom_types", "masses", "bonds", "angles", "dihedrals", "nonbonded"):
        g2["sys." + k] = system[k]
    g2["sys.cutoff"] = np.array(system["cutoff"])
    # the n54 statistics tables are dense over bead types, so they cover the 269-bead molecule's type tuples as well
    run_case("n269_b2", g2, extra={"pos": system["pos"], "atom_types": system["atom_types"],
                                   "cutoff": np.array(system["cutoff"])})
