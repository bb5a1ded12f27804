"""ORACLE support (test infrastructure): import the UNMODIFIED reference classes from
/root/reference inside the build container, to validate the restatements and to generate
the committed golden vectors (tests/golden/make_golden.py).  /root/reference does not
exist on the GPU box, so nothing at GPU-test / bench / smoke time may call this.

Shims:
  1. swin.py imports `open_clip` at module top (unused by Expert/MoE) -> stub modules.
  2. `import src.losses` pulls src/utils/__init__.py -> hydra/lightning; pre-register a bare
     namespace package `src.utils` so that __init__ never executes.
  3. losses.py uses `gather_tensor` without importing it (losses.py:512 vs :16) -> inject it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _install_shims() -> None:
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "open_clip" not in sys.modules:
        oc = types.ModuleType("open_clip")
        oct_ = types.ModuleType("open_clip.transformer")
        oct_.VisionTransformer = type("VisionTransformer", (), {})
        oc.transformer = oct_
        sys.modules["open_clip"] = oc
        sys.modules["open_clip.transformer"] = oct_
    if "src.utils" not in sys.modules:
        import src  # noqa: F401  (namespace package rooted at /root/reference/src)
        pkg = types.ModuleType("src.utils")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src", "utils")]
        sys.modules["src.utils"] = pkg


def load_moe_module():
    """-> module src.models.components.swin (Expert, MoE)."""
    _install_shims()
    return importlib.import_module("src.models.components.swin")


def load_losses_module():
    """-> module src.losses with the missing gather_tensor import patched in."""
    _install_shims()
    losses = importlib.import_module("src.losses")
    dist = importlib.import_module("src.utils.distributed")
    if not hasattr(losses, "gather_tensor"):
        losses.gather_tensor = dist.gather_tensor
    return losses
