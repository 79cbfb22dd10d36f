"""Import hook for the fixture generator: packages the reference imports at module level
but never reaches on the recorded code path (equinox / diffrax - pulse mode and the
batched jit path; matplotlib - drawing; optax, dill; jax sub-modules other than
jax.numpy / jax.random) resolve to empty placeholder modules, so that
`qml_essentials.model` / `ansaetze` / `gates` / `unitary` import unmodified.  Any attribute of
a placeholder is an inert object; a recorded case that really needed one of them would fail
loudly instead of producing numbers."""
import importlib.abc
import importlib.machinery
import sys
import types

ROOTS = {"diffrax", "matplotlib", "optax", "dill"}


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("placeholder object of a stubbed package was called")

    def __getattr__(self, name):
        return _Inert()

    def __mro_entries__(self, bases):
        return (object,)


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda name: _Inert()
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ROOTS or (
                name.startswith("jax.") and name not in ("jax.numpy", "jax.random")):
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
