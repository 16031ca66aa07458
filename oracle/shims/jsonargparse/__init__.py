"""Stand-in for the absent `jsonargparse` (only the names the reference imports)."""


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        raise RuntimeError("jsonargparse stand-in: CLI parsing is not available")


class ArgumentParser(_Dummy):
    pass


class ActionConfigFile(_Dummy):
    pass


class Namespace(dict):
    pass


def lazy_instance(cls, **kw):
    return cls(**kw)


def namespace_to_dict(ns):
    return dict(ns)


def class_from_function(fn, *a, **k):
    return fn
