"""`flashmd-nve-verlet` (reference scripts/nve_verlet.py)."""
from .nvt_langevin import run


def main(argv=None):
    from flashmd.simulation import NVESimulation
    return run(NVESimulation, "NVE velocity-Verlet simulation", argv=argv)


if __name__ == "__main__":
    main()
