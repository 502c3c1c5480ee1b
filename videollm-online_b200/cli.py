"""Headless driver + FPS measurement: the engine-backed counterpart of demo/cli.py:12-50.

    python -m videollm_online_b200.cli --synthetic_weights true [--frames 100] [--video clip.mp4]

Loads (or synthesises) a clip, asks for narration at t=0 and runs N iterations of
(encode 1 frame -> KV-append step -> maybe respond), reporting the reference's
"Average Processing FPS" = frames / wall time including responses (demo/cli.py:38)."""
from __future__ import annotations

import json
import os
import sys
import time

import torch

from .config import parse_args
from .inference import LiveInfer


def synthetic_clip(n_frames: int, resolution: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n_frames, 3, resolution, resolution), dtype=torch.uint8, generator=g)


def main(liveinfer: LiveInfer, video=None, n_iters: int = 100, save_history_path: str = None, quiet: bool = False):
    if video is None:
        video = synthetic_clip(n_iters, liveinfer.frame_resolution)
    liveinfer.load_video(video)
    liveinfer.input_query_stream('Please narrate the video in real time.', video_time=0.0)
    timecosts, history = [], {'frame_fps': liveinfer.frame_fps, 'conversation': []}
    fps = 0.0
    for i in range(min(n_iters, liveinfer.num_video_frames)):
        start = time.time()
        liveinfer.input_video_stream(i / liveinfer.frame_fps)
        query, response = liveinfer()
        timecosts.append(time.time() - start)
        fps = (i + 1) / sum(timecosts)
        entry = {'time': liveinfer.video_time, 'fps': fps, 'cost': timecosts[-1]}
        if query:
            history['conversation'].append({'role': 'user', 'content': query, **entry})
        if response:
            history['conversation'].append({'role': 'assistant', 'content': response, **entry})
        if not query and not response:
            history['conversation'].append(entry)
        if not quiet and (query or response):
            print(query or '', response or '')
    if save_history_path:
        os.makedirs(os.path.dirname(save_history_path) or '.', exist_ok=True)
        json.dump(history, open(save_history_path, 'w'), indent=4)
    if not quiet:
        print(f'Average Processing FPS: {fps:.1f}')
    return fps, history


if __name__ == '__main__':
    n, video = 100, None
    if '--frames' in sys.argv:
        n = int(sys.argv[sys.argv.index('--frames') + 1])
    if '--video' in sys.argv:   # a clip on disk: resampled to frame_fps and letterboxed to frame_resolution on load
        video = sys.argv[sys.argv.index('--video') + 1]
    main(LiveInfer(parse_args()), video=video, n_iters=n)
