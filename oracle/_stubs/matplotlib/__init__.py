"""Import stub (test infrastructure): the reference imports matplotlib at module import time
(src/rrt/rrt.py:9-11) but nothing on the hot path calls it; matplotlib is not installed here."""


def use(*args, **kwargs):
    return None


def __getattr__(name):
    raise AttributeError(f"matplotlib stub: {name} is not available")
