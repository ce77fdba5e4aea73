"""Import the reference's on-path modules from /root/reference WITHOUT running its
package __init__ (which pulls in the generate loop and fails on transformers >= 5;
SURVEY.md §8c).  /root/reference exists in the build container only; on the GPU box the loader
falls back to the verbatim copy under baseline/_ref (baseline/stage_reference.py), which is what
bench.py's reference arm times.  tests -m gpu and smoke() never call this.

Used by oracle/gen_golden.py (fixture generation) and by the optional
`tests/test_oracle_vs_reference.py` differential tests, which skip when the
reference checkout is absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")      # verbatim copy made by baseline/stage_reference.py (travels to the GPU box)
REF_ROOT = os.environ.get("SAMD_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "samd", "sam")) and os.path.isdir(os.path.join(STAGED_ROOT, "samd", "sam")):
    REF_ROOT = STAGED_ROOT


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "samd", "sam"))


def use_staged() -> bool:
    """Point the loader at baseline/_ref (bench.py's reference arm: the copy that exists on the GPU box too)."""
    global REF_ROOT
    if os.path.isdir(os.path.join(STAGED_ROOT, "samd", "sam")):
        REF_ROOT = STAGED_ROOT
        return True
    return False


def load():
    """Returns a namespace with the reference modules: .samd_sam, .samd_draft, .samd_utils,
    .samd_cache, .samd_config, .so_sam, .so_draft, .so_utils, .so_config, .tr_utils."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)          # profile_utils lives at the reference root
    for pkg in ("samd", "samd_sam_only"):
        mod = sys.modules.get(pkg)
        if mod is None or not getattr(mod, "__ref_stub__", False):
            stub = types.ModuleType(pkg)
            stub.__path__ = [os.path.join(REF_ROOT, pkg)]
            stub.__ref_stub__ = True
            sys.modules[pkg] = stub
            for k in [k for k in sys.modules if k.startswith(pkg + ".")]:
                del sys.modules[k]
    ns = types.SimpleNamespace()
    ns.samd_config = importlib.import_module("samd.samd_config")
    ns.samd_sam = importlib.import_module("samd.sam")
    ns.samd_draft = importlib.import_module("samd.draft")
    ns.samd_utils = importlib.import_module("samd.utils")
    ns.samd_cache = importlib.import_module("samd.cache")
    ns.tr_utils = importlib.import_module("samd.tree_model.token_recycle.utils")
    ns.so_config = importlib.import_module("samd_sam_only.samd_config")
    ns.so_sam = importlib.import_module("samd_sam_only.sam")
    ns.so_draft = importlib.import_module("samd_sam_only.draft")
    ns.so_utils = importlib.import_module("samd_sam_only.utils")
    return ns


def unload():
    """Drop the stub packages so the repo's own samd / samd_sam_only can be imported."""
    for k in [k for k in sys.modules if k == "samd" or k == "samd_sam_only" or k.startswith("samd.")
              or k.startswith("samd_sam_only.") or k == "profile_utils"]:
        del sys.modules[k]
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
