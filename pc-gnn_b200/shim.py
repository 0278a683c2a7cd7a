"""Drop-in wiring: make the reference's own trainer import the B200 path.

The reference's ``src/model_handler.py`` does

    from src.utils import test, load_data, pos_neg_split, normalize, pick_step, set_seeds
    from src.model import PCALayer
    from src.layers import InterAgg1, InterAgg3, InterAgg5, IntraAgg
    from src.graphsage import *

(/root/reference/src/model_handler.py:10-14). ``install()`` registers this package's modules under
those names in ``sys.modules`` so that the trainer, the reference's ``src/model.py`` and
``src/result_manager.py`` run unchanged on top of the CUDA kernels:

    import pcgnn_b200.shim as shim
    shim.install(reference_root="/path/to/PC-GNN")      # keeps src.model / src.result_manager from there
    from src.model_handler import ModelHandler

Modules the hot path does not touch (``src.model``, ``src.result_manager``, ``src.model_handler``) are
still loaded from the reference tree when ``reference_root`` is given; without it ``src.model`` falls
back to this package's mirror.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

__all__ = ["install", "uninstall"]

_REPLACED = ("src.layers", "src.graphsage", "src.utils")


def install(reference_root: str | None = None, replace_model: bool = False):
    from . import graphsage, layers, model, utils

    pkg = types.ModuleType("src")
    pkg.__path__ = [os.path.join(reference_root, "src")] if reference_root else []
    pkg.__package__ = "src"
    sys.modules["src"] = pkg
    sys.modules["src.layers"] = layers
    sys.modules["src.graphsage"] = graphsage
    sys.modules["src.utils"] = utils
    pkg.layers, pkg.graphsage, pkg.utils = layers, graphsage, utils
    if replace_model or not reference_root:
        sys.modules["src.model"] = model
        pkg.model = model
    return pkg


def uninstall():
    for name in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[name]
