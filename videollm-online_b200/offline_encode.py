"""Batched offline feature extraction — the host loop of `distributed_encode`
(data/utils.py:86-104, driven by data/preprocess/encode.py:19-27) over an arbitrary `vision_encode` callable.

For every clip of a directory: decode -> split into batches -> `vision_encode(encoder, batch)` -> concatenate the
[T, frame_num_tokens, vision_hidden] tokens -> optional bf16 -> `torch.save` under
`<src_root>_<embed_mark tail>_<vision_pretrained with '/' -> '--'>/<clip>.pt` (the layout data/stream.py:90-91 loads).
Clips are assigned to ranks round-robin by directory-listing index, as the reference does with submitit's
(global_rank, num_tasks); here the pair comes from torch.distributed / the caller.

With the engine, `encoder, vision_encode = build_live_vision(config, engine)`: the ViT kernels run at the batch size
given here (the engine chunks by its `max_vit_batch`), the tensor-core-bound regime of the same kernels the
streaming path uses at batch 1.  Pure host code: no arithmetic of its own.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch


def encoded_root(src_root: str, embed_mark: str, vision_pretrained: str) -> str:
    src_root = src_root.rstrip('/')
    return f"{src_root}_{embed_mark.split('_')[-1]}_{vision_pretrained.replace('/', '--')}"


def encode_directory(*, src_root: str, vision_pretrained: str, vision_encode: Callable, encoder=None, batch_size: int = 256,
                     embed_mark: str = "2fps_384_1+3x3", save_bf16: bool = False, rank: int = 0, world_size: int = 1,
                     device: Optional[str] = None, read_video: Optional[Callable] = None) -> List[str]:
    """Returns the `.pt` paths this rank wrote."""
    if read_video is None:
        from .video_ingest import read_video_resampled as read_video
    src_root = src_root.rstrip('/')
    dst_root = encoded_root(src_root, embed_mark, vision_pretrained)
    os.makedirs(dst_root, exist_ok=True)
    written = []
    for i, file in enumerate(sorted(os.listdir(src_root))):   # sorted: every rank must see the same order
        if i % world_size != rank:
            continue
        frame_path = os.path.join(src_root, file)
        if not os.path.isfile(frame_path):
            continue
        save_path = (os.path.splitext(frame_path)[0] + '.pt').replace(src_root, dst_root, 1)
        frames = read_video(frame_path)                       # uint8 [T, 3, H, W], already at the target fps / size
        with torch.no_grad():
            feats = torch.cat([vision_encode(encoder, batch.to(device) if device else batch).cpu()
                               for batch in frames.split(batch_size)])
        if save_bf16:
            feats = feats.to(torch.bfloat16)
        torch.save(feats, save_path)
        written.append(save_path)
    return written
