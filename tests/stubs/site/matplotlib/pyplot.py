from _ub_dummy import Dummy


def subplots(*args, **kwargs):
    return Dummy(), Dummy()


def ylim(*args, **kwargs):
    return (0.0, 1.0)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Dummy()
