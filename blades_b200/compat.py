"""Drop-in compatibility for scripts written against the reference package.

``import blades_b200.compat; blades_b200.compat.install()`` (or ``python -m blades_b200.compat script.py [args]``)
makes ``import blades``, ``from blades.simulator import Simulator``, ``from blades.datasets import MNIST``,
``from blades.models.mnist import MLP``, ``from blades.client import ByzantineClient``, ... resolve to the
corresponding ``blades_b200`` modules (same module objects, no copies), so the reference's examples and user scripts
(reference ``src/blades/examples/*.py``, ``scripts/*.py``) run unchanged.  Those scripts also call ``ray.init(...)``:
when Ray is not installed a no-op stand-in is registered -- this package never uses Ray (one process per GPU under
``torchrun`` replaces the actors), so ``ray.init`` has nothing to do.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

__all__ = ["install", "uninstall", "run_script"]

_ALIAS, _REAL = "blades", "blades_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``blades[.x.y]`` -> the already importable ``blades_b200[.x.y]`` module object."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != _ALIAS and not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            mod = importlib.import_module(real)
        except ImportError:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=hasattr(mod, "__path__"))

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):          # the real module is already initialised
        pass


def _ray_stub() -> types.ModuleType:
    ray = types.ModuleType("ray")
    ray.__doc__ = "no-op stand-in registered by blades_b200.compat (Ray is not used by blades_b200)"
    ray.init = lambda *a, **k: None
    ray.shutdown = lambda *a, **k: None
    ray.is_initialized = lambda: True
    return ray


_finder = None


def install(ray_stub: bool = True) -> None:
    """Register the ``blades`` alias (idempotent).  ``ray_stub``: also register a no-op ``ray`` if Ray is missing."""
    global _finder
    if _finder is None:
        if _ALIAS in sys.modules and not getattr(sys.modules[_ALIAS], "__name__", "").startswith(_REAL):
            raise RuntimeError("a different 'blades' package is already imported in this process")
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)
    if ray_stub and "ray" not in sys.modules:
        try:
            importlib.import_module("ray")
        except ImportError:
            sys.modules["ray"] = _ray_stub()


def uninstall() -> None:
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in [n for n in sys.modules if n == _ALIAS or n.startswith(_ALIAS + ".")]:
        del sys.modules[name]
    if getattr(sys.modules.get("ray"), "__doc__", "") and "blades_b200.compat" in sys.modules["ray"].__doc__:
        del sys.modules["ray"]


def run_script(path: str, argv=None, patch=None) -> dict:
    """Execute a reference-style script with the alias installed; returns its globals.  ``patch`` (str -> str) may
    rewrite the source first (tests shorten the hard-coded round counts this way)."""
    import runpy
    install()
    old_argv = sys.argv
    sys.argv = [path] + list(argv or [])
    try:
        if patch is None:
            return runpy.run_path(path, run_name="__main__")
        src = patch(open(path).read())
        glob = {"__name__": "__main__", "__file__": path}
        exec(compile(src, path, "exec"), glob)          # noqa: S102 - the caller's own script
        return glob
    finally:
        sys.argv = old_argv


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m blades_b200.compat script.py [args...]")
    run_script(sys.argv[1], sys.argv[2:])
