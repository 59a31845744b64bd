"""Import stub (test infrastructure); see matplotlib/__init__.py."""


def __getattr__(name):
    def _missing(*args, **kwargs):
        raise RuntimeError(f"matplotlib stub: {name} called")
    return _missing
