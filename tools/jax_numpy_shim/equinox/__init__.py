"""equinox stand-in for the fixture generator: `filter_jit` is the identity (eager NumPy),
`Module` a plain base class.  See ../jax/__init__.py."""


class Module:
    pass


def filter_jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn
