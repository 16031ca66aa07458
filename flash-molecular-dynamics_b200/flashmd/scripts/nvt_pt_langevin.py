"""`flashmd-pt-langevin` (reference scripts/nvt_pt_langevin.py:31-58)."""
from .nvt_langevin import run


def main(argv=None):
    from flashmd.simulation import PTSimulation
    return run(PTSimulation, "Parallel-tempering Langevin NVT simulation", betas_are_list=True, argv=argv)


if __name__ == "__main__":
    main()
