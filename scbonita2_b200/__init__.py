"""scbonita2_b200 — the data-parallel hot path of scBONITA2 on B200 (sm_100a).

Host side: Python modules that mirror the reference's entry points (``rule_determination``,
``importance_scores``, ``attractor_analysis``, ``node_class`` ...).  Device side: hand-written
CUDA kernels behind the C-ABI of ``include/scbonita_b200.h`` (``csrc/libscbonita_b200.so``).
There is no CPU fallback: anything that computes needs the library and a CUDA device.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

# reference top-level module name -> our module (the reference runs its scripts from code/,
# so its pickles name classes as e.g. ``node_class.Node``)
_ALIASES = ("node_class", "network_class", "cell_class", "file_paths", "rule_determination",
            "importance_scores", "attractor_analysis", "rule_inference", "pipeline_class")
_PICKLED = {"node_class": ("Node",), "network_class": ("Network",),
            "cell_class": ("Cell", "CellPopulation"), "rule_inference": ("RuleInference",)}


def install_aliases(force: bool = False) -> None:
    """Make ``import node_class`` (etc.) resolve to this package and pickle our model classes
    under the reference's module paths, so ``*.ruleset.pickle`` / ``*.network.pickle`` /
    ``*.cells.pickle`` files are interchangeable with the reference's
    (``pipeline_class.py:160-182``, ``importance_scores.py:386-415``)."""
    for name in _ALIASES:
        try:
            module = importlib.import_module(f"{__name__}.{name}")
        except ModuleNotFoundError:
            continue
        if force or name not in sys.modules:
            sys.modules[name] = module
        if sys.modules.get(name) is module:
            for cls_name in _PICKLED.get(name, ()):
                getattr(module, cls_name).__module__ = name
