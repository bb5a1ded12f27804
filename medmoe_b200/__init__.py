"""medmoe_b200 — B200-native (sm_100a) MoE block + global contrastive loss of MedMoE.

Public surface mirrors the reference's operator interface for this path
(src/models/components/swin.py `MoE`, src/losses.py global contrastive losses,
src/utils/distributed.py gather helpers); see INTEGRATION.md.
"""
__version__ = "0.1.0"
