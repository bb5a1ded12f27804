"""Swin-T backbone + MoE block, the drop-in for the reference's `SWIN` wrapper (src/models/components/swin.py:120-151).

Same constructor arguments, same `forward(x) -> (global_feat, local_feat, router_probs)`, same attribute names
(`model`, `moe`) so state dicts line up.  What changes around the MoE:
  * the reference runs the HF `AutoImageProcessor` (PIL, CPU, per step) inside `forward` (swin.py:131); here the same
    resize / rescale / ImageNet normalisation is a few tensor ops on the GPU (`preprocess`), and can be switched off when the
    data pipeline already delivers normalised 224x224 tensors;
  * the backbone runs under bf16 autocast and its four stage outputs are handed to the MoE as they are (token-major
    [B, P_s, D_s], the layout the dispatch kernel reads), no copies, no dtype round trip;
  * `swin_feat` (router input) = mean over the tokens of the last hidden state, in fp32.
The Swin backbone itself is HuggingFace `transformers` code (pinned 4.46.0 by the reference's environment.yml) and outside the
parity boundary; there is no network here, so `pretrained=True` only works with a local HF cache.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .moe import MoE

SWIN_MODEL_ID = "microsoft/swin-tiny-patch4-window7-224"
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class SWIN(nn.Module):
    def __init__(self, pretrained: bool = True, lora: bool = False, lora_r: int = 8, lora_alpha: int = 16,
                 lora_dropout: float = 0.1, use_moe: bool = True, *, num_experts: int = 6, topk: int = 1,
                 image_size: int = 224, preprocess: bool = True, model_id: str = SWIN_MODEL_ID):
        super().__init__()
        if lora:
            raise NotImplementedError("LoRA adapters (peft) are outside the B200 hot path; load merged weights instead")
        from transformers import SwinConfig, SwinModel   # heavy import, only when a backbone is built

        self.moe: Optional[MoE] = MoE(num_experts=num_experts, topk=topk) if use_moe else None
        if pretrained:
            self.model = SwinModel.from_pretrained(model_id, local_files_only=True)
        else:
            self.model = SwinModel(SwinConfig(image_size=image_size))
        self.image_size = int(self.model.config.image_size)
        self.do_preprocess = bool(preprocess)
        self.register_buffer("pixel_mean", torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(IMAGENET_STD).view(1, 3, 1, 1), persistent=False)

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        """uint8 [0, 255] or float [0, 1] images [B, 3, H, W] -> normalised [B, 3, S, S] (what AutoImageProcessor produces)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("SWIN expects images as a [B, 3, H, W] tensor")
        x = x.float() / 255.0 if x.dtype == torch.uint8 else x.float()
        if x.shape[-2:] != (self.image_size, self.image_size):
            x = F.interpolate(x, size=(self.image_size, self.image_size), mode="bicubic", align_corners=False, antialias=True)
        return (x - self.pixel_mean) / self.pixel_std

    def stage_features(self, pixel_values: torch.Tensor):
        """(the four stage outputs [B, P_s, D_s], swin_feat [B, 768] fp32, last hidden state)"""
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=pixel_values.is_cuda):
            out = self.model(pixel_values=pixel_values, output_hidden_states=True)
        final_hidden = out.last_hidden_state
        return [out.hidden_states[i] for i in range(4)], final_hidden.float().mean(dim=1), final_hidden

    def forward(self, x: torch.Tensor):
        pixel_values = self.preprocess(x) if self.do_preprocess else x
        stage_feats, swin_feat, final_hidden = self.stage_features(pixel_values)
        if self.moe is None:                         # swin.py:145-148
            return final_hidden.mean(dim=1), final_hidden, None
        return self.moe(stage_feats, swin_feat)
