def __getattr__(name):
    from _ub_dummy import module_getattr
    return module_getattr(name)
