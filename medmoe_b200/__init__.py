"""medmoe_b200 — B200-native (sm_100a) MoE block + global contrastive loss of MedMoE.

Public surface mirrors the reference's operator interface for this path
(src/models/components/swin.py `MoE`/`Expert`, src/losses.py global contrastive losses,
src/utils/distributed.py gather helpers); see INTEGRATION.md for the drop-in recipe.
"""
from .distributed import BackpropType, OverlappedGradSync, concat_gather_all_gpu, gather_tensor, get_rank
from .losses import (ContrastiveLossOutput, FLAVAGlobalContrastiveLoss, FLAVAGlobalContrastiveLossOutput,
                     GLORIAGlobalContrastiveLoss, contrastive_loss_with_temperature, zero_shot_predict)
from .moe import Expert, MoE
from .local_loss import GLORIALocalContrastiveLoss, GLORIALocalContrastiveLossOutput, local_similarities
from .checkpoint import extract_moe_state_dict, load_reference_checkpoint
from .eval_zs import zero_shot_evaluate

__version__ = "0.1.0"

__all__ = [
    "MoE", "Expert", "GLORIAGlobalContrastiveLoss", "FLAVAGlobalContrastiveLoss", "FLAVAGlobalContrastiveLossOutput",
    "ContrastiveLossOutput", "contrastive_loss_with_temperature", "zero_shot_predict", "BackpropType", "gather_tensor",
    "concat_gather_all_gpu", "get_rank", "activate", "load_reference_checkpoint", "extract_moe_state_dict",
    "zero_shot_evaluate", "SWIN", "OverlappedGradSync", "GLORIALocalContrastiveLoss", "GLORIALocalContrastiveLossOutput", "local_similarities",
]


def __getattr__(name):
    if name == "SWIN":            # lazy: pulls in `transformers`
        from .swin import SWIN
        return SWIN
    raise AttributeError(name)


def activate():
    """Swap the B200 modules into an importable reference checkout (the whole integration):
    `src.models.components.swin.MoE` (looked up as a module global by SWIN.__init__, swin.py:123)
    and the loss classes named by Hydra `_target_` strings (configs/model/med-moe_pretraining.yaml:31)."""
    import importlib

    swin = importlib.import_module("src.models.components.swin")
    swin.MoE, swin.Expert = MoE, Expert
    losses = importlib.import_module("src.losses")
    losses.GLORIAGlobalContrastiveLoss = GLORIAGlobalContrastiveLoss
    losses.GLORIALocalContrastiveLoss = GLORIALocalContrastiveLoss
    losses.FLAVAGlobalContrastiveLoss = FLAVAGlobalContrastiveLoss
    losses.contrastive_loss_with_temperature = contrastive_loss_with_temperature
    return swin, losses
