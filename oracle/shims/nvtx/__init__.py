"""Stand-in for the absent `nvtx` package (test infrastructure only).

`annotate` works both as a decorator and as a context manager and does nothing.
"""
import functools


class annotate:
    def __init__(self, message=None, color=None, domain=None, category=None, payload=None):
        self.message = message

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __call__(self, fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            return fn(*a, **k)

        return wrapped


def push_range(*a, **k):
    return None


def pop_range(*a, **k):
    return None
