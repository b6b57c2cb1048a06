"""``import pgl`` -> pgl_b200 (the drop-in keeps PGL's import surface).

A thin alias, no code of its own: ``pgl`` re-exports pgl_b200's public names and a meta-path finder maps every
``pgl.<sub>`` import (pgl.nn, pgl.nn.functional, pgl.utils.op, pgl.math, pgl.partition, pgl.sampling ...) to the
``pgl_b200.<sub>`` module object itself, so ``pgl.graph.Graph is pgl_b200.graph.Graph``.

Only usable where the reference itself is not installed under the same name (on a box that has PaddlePaddle/PGL,
import ``pgl_b200`` explicitly)."""
import importlib
import importlib.abc
import importlib.util
import sys

import pgl_b200 as _impl
from pgl_b200 import *  # noqa: F401,F403


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("pgl."):
            return None
        real = "pgl_b200." + fullname[len("pgl."):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


sys.meta_path.insert(0, _AliasFinder())
for _name in dir(_impl):
    if not _name.startswith("__"):
        globals().setdefault(_name, getattr(_impl, _name))
__version__ = getattr(_impl, "__version__", "0")
