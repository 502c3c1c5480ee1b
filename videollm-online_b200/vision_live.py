"""`build_live_vision` over the engine (reference: models/vision_live.py:10-30,54-61).

`encode_fn(encoder, frames)` returns the [B, frame_num_tokens, vision_hidden] tokens — CLS := MAP-head
pooler_output followed by the row-major 3x3 adaptive-average-pooled patch tokens — which is also what
the offline feature extraction consumes (data/preprocess/encode.py:19, data/utils.py:86-104)."""
from __future__ import annotations

from functools import partial

import torch

from .engine import Engine, VloError


class VisionEncoder:
    """Handle standing in for `AutoModel.from_pretrained(...).vision_model`."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def to(self, *a, **k):
        return self

    def eval(self):
        return self


def _siglip_vision_encode(vision_model: VisionEncoder, frames: torch.Tensor, frame_token_cls: bool = True,
                          frame_token_pooled=(3, 3), **kwargs) -> torch.Tensor:
    cfg = vision_model.engine.cfg
    if bool(frame_token_cls) != bool(cfg.frame_token_cls) or list(frame_token_pooled or []) != list(cfg.frame_token_pooled or []):
        raise VloError("the engine was built for a different frame-token layout")
    _, tokens = vision_model.engine.vit_encode(frames, return_vit_tokens=True, connector=False)
    return tokens


def build_live_vision(config, engine: Engine = None):
    if engine is None:
        raise VloError("build_live_vision needs the engine that holds the vision tower")
    if "siglip" not in (config.vision_pretrained or "siglip"):
        raise ValueError(f"Unverified vision_pretrained: {config.vision_pretrained}")
    return VisionEncoder(engine), partial(_siglip_vision_encode, frame_token_cls=config.frame_token_cls,
                                          frame_token_pooled=config.frame_token_pooled)
