"""`flashmd-langevin` (reference scripts/nvt_langevin.py:6-179).  `--disable_optim` must act before
`flashmd.models.schnet` is imported because the MLCG_* toggles are read at import time."""
import os
import sys

DISABLE_OPTIM = "--disable_optim" in sys.argv
if DISABLE_OPTIM:
    sys.argv.remove("--disable_optim")
    for _k in ("MLCG_USE_TRITON_MESSAGE_PASSING", "MLCG_USE_FUSED_RBF", "MLCG_USE_FUSED_TANH_LINEAR", "MLCG_USE_CSR",
               "MLCG_USE_SRC_CSR_GRAD_X"):
        os.environ[_k] = "0"


def run(simulation_class, description, betas_are_list=False, argv=None):
    import json
    import torch
    from flashmd.simulation.cli import parse_simulation_config
    torch.set_float32_matmul_precision("high")
    model, data_list, betas, sim, profile = parse_simulation_config(simulation_class, description, argv=argv)
    if DISABLE_OPTIM:
        sim.gptq = None                 # plain fp32 modules, no fused engine: the reference's CPU-capable path
        sim._compile_model_flag = False
        sim.force_module_path = True
    if betas_are_list and not isinstance(betas, list):
        betas = [betas]
    if betas_are_list:
        sim.attach_model_and_configurations(model, data_list, betas=betas)
    else:
        sim.attach_model_and_configurations(model, data_list, beta=betas)
    if profile:
        os.makedirs(profile, exist_ok=True)
        with torch.profiler.profile(on_trace_ready=torch.profiler.tensorboard_trace_handler(profile)) as prof:
            sim.simulate(prof=prof)
    else:
        sim.simulate()
    m = sim.get_throughput_metrics()
    if m:
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in m.items()}))
    return sim


def main(argv=None):
    from flashmd.simulation import LangevinSimulation
    return run(LangevinSimulation, "Langevin (BAOAB) NVT simulation", argv=argv)


if __name__ == "__main__":
    main()
