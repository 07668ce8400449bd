"""Stand-in for matplotlib (plots are presentation, out of scope) -- TEST INFRASTRUCTURE ONLY.  Lets the reference's
scoring functions, which draw figures in between the arithmetic (eval_uncertainty.py:465-550), run headless."""


def use(*args, **kwargs):
    return None


def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
