"""Import the REAL scBONITA2 reference (``/root/reference/code``) inside the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``scbonita2_b200``) imports this
module.  It exists so that (a) the CPU restatements under ``oracle/`` can be pinned against
the unmodified reference and (b) ``oracle/make_golden.py`` can generate the committed
fixtures in ``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so every
user of this module must skip cleanly when :func:`available` is False.

The reference writes its outputs relative to its own source directory
(``file_paths.py:4-17``) and ``/root/reference`` is read-only, therefore the sources are
copied to a throw-away temp dir *outside the repo* at import time (never into the repo).
Third-party modules the container lacks are replaced by minimal in-memory stand-ins:

* ``alive_progress.alive_bar``        -> no-op context manager (progress bar only)
* ``numexpr.evaluate``                -> ``eval`` on NumPy arrays; the reference only feeds it
  ``& | ~`` over bool arrays (``importance_scores.py:50-57``, ``attractor_analysis.py:52-58``)
  whose semantics are identical
* ``matplotlib`` / ``seaborn`` / ``bioservices`` / ``bs4`` / ``fastdtw`` / ``kneed`` -> empty
  stubs (imported at module top, never reached on the hot path)
"""
from __future__ import annotations

import atexit
import contextlib
import importlib
import os
import shutil
import sys
import tempfile
import types

REFERENCE_CODE_DIR = "/root/reference/code"
# git-ignored install of the unmodified reference sources (``install_baseline`` below, run by
# ``__graft_entry__.build()`` in the build container): unlike /root/reference it travels to the GPU
# box, where ``bench.py --impl reference`` and the sweep's ``cpu_baseline`` time it
BASELINE_CODE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "code")
_STATE: dict = {}


def code_dir():
    for cand in (REFERENCE_CODE_DIR, BASELINE_CODE_DIR):
        if os.path.isfile(os.path.join(cand, "importance_scores.py")):
            return cand
    return None


def available() -> bool:
    return code_dir() is not None


def install_baseline() -> bool:
    """Copy the reference's ``code/*.py`` (unmodified) into git-ignored ``baseline/_ref/code``.
    The reference is plain Python without a setup script, so this copy IS its install.  Returns
    whether the install exists afterwards."""
    if os.path.isdir(REFERENCE_CODE_DIR):
        os.makedirs(BASELINE_CODE_DIR, exist_ok=True)
        for name in sorted(os.listdir(REFERENCE_CODE_DIR)):
            if name.endswith(".py"):
                shutil.copyfile(os.path.join(REFERENCE_CODE_DIR, name), os.path.join(BASELINE_CODE_DIR, name))
    return os.path.isfile(os.path.join(BASELINE_CODE_DIR, "importance_scores.py"))


class _Anything:
    """Callable/attribute sink used for plotting & network stubs."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):          # "fig, ax = plt.subplots(...)"
        return iter((_Anything(), _Anything()))


def _stub_module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr: _Anything()  # type: ignore[attr-defined]
    sys.modules[name] = mod
    return mod


def _install_shims() -> None:
    # real packages that probe for optional deps must be imported BEFORE the stand-ins exist
    for real in ("pandas", "sklearn", "sklearn.preprocessing", "scipy.stats", "networkx"):
        try:
            importlib.import_module(real)
        except Exception:
            pass
    @contextlib.contextmanager
    def alive_bar(*_a, **_k):
        yield lambda *a, **k: None

    if "alive_progress" not in sys.modules:
        _stub_module("alive_progress", alive_bar=alive_bar)

    def ne_evaluate(expression, local_dict=None, global_dict=None):
        return eval(expression, {"__builtins__": {}}, dict(local_dict or {}))

    try:
        import numexpr  # noqa: F401
    except Exception:
        _stub_module("numexpr", evaluate=ne_evaluate)

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches",
                 "seaborn", "bioservices", "bioservices.kegg", "bs4", "fastdtw", "kneed"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub_module(name)
    # package attribute wiring for "import matplotlib.pyplot as plt"
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and \
            not hasattr(sys.modules["matplotlib"], "__path__"):
        sys.modules["matplotlib"].__path__ = []  # type: ignore[attr-defined]
        sys.modules["bioservices"].__path__ = []  # type: ignore[attr-defined]


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference modules (imported once per process)."""
    if "ns" in _STATE:
        return _STATE["ns"]
    src = code_dir()
    if src is None:
        raise RuntimeError("reference sources are neither mounted at " + REFERENCE_CODE_DIR +
                           " nor installed at " + BASELINE_CODE_DIR)
    _install_shims()
    root = tempfile.mkdtemp(prefix="scbonita_ref_")
    atexit.register(shutil.rmtree, root, ignore_errors=True)
    code = os.path.join(root, "code")
    shutil.copytree(src, code, ignore=shutil.ignore_patterns("__pycache__"))
    sys.path.insert(0, code)
    ns = types.SimpleNamespace(root=root, code_dir=code, source=src)
    for name in ("file_paths", "node_class", "network_class", "cell_class",
                 "rule_determination", "importance_scores", "rule_inference"):
        # a product alias of the same top-level name must not shadow the reference
        sys.modules.pop(name, None)
        setattr(ns, name, importlib.import_module(name))
    try:
        sys.modules.pop("attractor_analysis", None)
        ns.attractor_analysis = importlib.import_module("attractor_analysis")
    except Exception as exc:  # pragma: no cover - depends on container packages
        ns.attractor_analysis = None
        ns.attractor_analysis_error = exc
    _STATE["ns"] = ns
    return ns
