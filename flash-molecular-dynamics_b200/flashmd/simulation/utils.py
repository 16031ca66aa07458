import yaml

KB_KCAL = 0.0019872041   # kcal/mol/K


KBOLTZMANN = 1.38064852e-23   # J/K          (reference simulation/base.py:36-38)
AVOGADRO = 6.022140857e23
JPERKCAL = 4184


def calc_beta_from_temperature(temp):
    """Temperature(s) in Kelvin -> inverse temperature(s) in mol/kcal (reference simulation/utils.py)."""
    import numpy as np
    return JPERKCAL / KBOLTZMANN / AVOGADRO / np.array(temp)


def beta_from_temperature(temperature_K: float) -> float:
    return 1.0 / (KB_KCAL * temperature_K)


def temperature_from_beta(beta: float) -> float:
    return 1.0 / (KB_KCAL * beta)


def load_yaml(path: str) -> dict:
    with open(path) as fh:
        return yaml.safe_load(fh) or {}


def dump_yaml(path: str, obj: dict) -> None:
    with open(path, "w") as fh:
        yaml.safe_dump(obj, fh, sort_keys=False)
