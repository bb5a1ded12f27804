"""Reference checkpoints -> medmoe_b200.MoE.

The MoE block keeps the reference's parameter names (`experts.{e}.proj_convs.{s}.0.weight`, `experts.{e}.attn_proj.{0,2}.*`,
`router.{0,2}.*`, swin.py:18-30,83-92), so loading is a matter of finding the block inside a larger state dict:
a Lightning checkpoint of `MedMoELitModule` stores it under `model.image_encoder.model.moe.`
(medmoe_module.py:69 -> med_moe.py:34 -> vision_encoder.py:21 -> swin.py:123), an `ImageEncoder` state dict under
`model.moe.`, a bare `SWIN` under `moe.`.
"""
from __future__ import annotations

import re
from typing import Dict, Mapping, Optional, Tuple

import torch

REFERENCE_PREFIXES: Tuple[str, ...] = ("model.image_encoder.model.moe.", "image_encoder.model.moe.", "model.moe.", "moe.", "")
_MOE_KEY = re.compile(r"^(experts\.\d+\.(proj_convs\.\d+\.0|attn_proj\.[02])|router\.[02])\.(weight|bias)$")


def rename_medclip_vision_keys(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference's MedCLIP rename for the vision tower (med_moe.py:42-44): keep `vision_model.*`, call it `model.*`."""
    return {k.replace("vision_model.", "model."): v for k, v in state_dict.items() if "vision_model" in k}


def extract_moe_state_dict(state: Mapping, prefix: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Pick the MoE block's tensors out of a checkpoint / state dict and strip the prefix.

    `state` may be a Lightning checkpoint (has a "state_dict" entry) or a plain state dict.  With `prefix=None` the known
    reference prefixes are tried, longest first; raises KeyError when no MoE parameters are found."""
    sd = state["state_dict"] if "state_dict" in state and isinstance(state["state_dict"], Mapping) else state
    for pre in (REFERENCE_PREFIXES if prefix is None else (prefix,)):
        picked = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre) and _MOE_KEY.match(k[len(pre):])}
        if picked:
            return picked
    raise KeyError("no MoE parameters (experts.*, router.*) found under " +
                   (repr(prefix) if prefix is not None else "any of " + ", ".join(map(repr, REFERENCE_PREFIXES))))


def num_experts_in(moe_state: Mapping[str, torch.Tensor]) -> int:
    ids = {int(k.split(".")[1]) for k in moe_state if k.startswith("experts.")}
    return max(ids) + 1 if ids else 0


def load_reference_checkpoint(moe: torch.nn.Module, source, prefix: Optional[str] = None, strict: bool = True):
    """Load the MoE weights of a reference checkpoint (path or already-loaded mapping) into a medmoe_b200.MoE.

    Shapes must match the module (number of experts, hidden dims); a mismatch raises like `load_state_dict` does.
    Returns the `load_state_dict` result.  The bf16 shadow copies the kernels read are refreshed on the next forward."""
    if not isinstance(source, Mapping):
        source = torch.load(source, map_location="cpu", weights_only=True)
    moe_state = extract_moe_state_dict(source, prefix)
    want = len(moe.experts)
    have = num_experts_in(moe_state)
    if have != want:
        raise RuntimeError(f"checkpoint holds {have} experts, the module was built with num_experts={want}")
    return moe.load_state_dict(moe_state, strict=strict)
