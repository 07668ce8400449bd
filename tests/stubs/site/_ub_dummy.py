"""Permissive dummy objects shared by the stub packages -- TEST INFRASTRUCTURE ONLY."""


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make(name)

    def __getitem__(cls, item):
        return cls


class Dummy(metaclass=_Meta):
    """Accepts anything: construction with any arguments, attribute access, calls, decoration, iteration."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs
        for k, v in kwargs.items():
            try:
                object.__setattr__(self, k, v)
            except Exception:
                pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Dummy()

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]           # used as a decorator
        return Dummy()

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return Dummy()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __truediv__(self, other):
        return self


def _make(name):
    return _Meta(name, (Dummy,), {})


def module_getattr(name):
    """``__getattr__`` for explicit stub modules: fabricate any name they do not define."""
    if name.startswith("__"):
        raise AttributeError(name)
    return _make(name)
