import yaml

KB_KCAL = 0.0019872041   # kcal/mol/K


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
