"""Video ingest for the streaming demo: decode -> frame-rate resample -> letterbox to a square.

The reference preprocesses a clip once with an external ffmpeg binary (`ffmpeg_once`, data/utils.py:51-66, called
from demo/cli.py:15-20: `-r <fps>` and `scale` of the longer side to `resolution` followed by a centred `pad` to
`resolution x resolution` in black) and then decodes the cached file with `torchvision.io.read_video`
(demo/inference.py:111-115).  Neither the ffmpeg binary nor a torchvision video decoder exists in this image, so the
same geometry and sampling rule are applied here in one pass with OpenCV:

  * output frame k shows the source frame whose timestamp is nearest to k / fps (ffmpeg's fps conversion with its
    default rounding), for k < round(duration * fps);
  * the longer side is scaled to `resolution` with a bicubic filter (ffmpeg_once's `-sws_flags bicubic`), the shorter
    side to the nearest EVEN size that keeps the aspect ratio (ffmpeg's `-2`), and the result is centred on a black
    square ((ow-iw)/2, (oh-ih)/2 rounded down).

Not bit-identical to libswscale's bicubic taps (documented; the frames then go through the same uint8 -> ViT path).
Host-side only: returns a uint8 [T, 3, R, R] CPU tensor in RGB that `LiveInfer.load_video` moves to the GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def letterbox_geometry(width: int, height: int, resolution: int) -> Tuple[int, int, int, int]:
    """(scaled_w, scaled_h, x0, y0) of ffmpeg_once's scale + pad filter chain."""
    if width > height:
        sw = resolution
        sh = int(round(height * resolution / width / 2.0)) * 2
    else:
        sh = resolution
        sw = int(round(width * resolution / height / 2.0)) * 2
    sw, sh = max(2, min(sw, resolution)), max(2, min(sh, resolution))
    return sw, sh, (resolution - sw) // 2, (resolution - sh) // 2


def resample_indices(n_src: int, src_fps: float, fps: float) -> list:
    """Source frame index shown by every output frame of a constant-frame-rate conversion to `fps`."""
    if n_src <= 0 or src_fps <= 0 or fps <= 0:
        return []
    n_out = max(1, int(round(n_src / src_fps * fps)))
    return [min(n_src - 1, int(round(k / fps * src_fps))) for k in range(n_out)]


def read_video_resampled(path: str, fps: Optional[float] = None, resolution: Optional[int] = None) -> torch.Tensor:
    """Decode `path`; optionally convert to `fps` frames/s and letterbox to `resolution`.  uint8 [T,3,H,W], RGB.
    Streaming: only the frames the output keeps are converted and held (a 107 s 1440x1080 clip is 3204 source frames =
    15 GB decoded, but 214 frames of 384x384 at 2 FPS)."""
    import numpy as np
    try:
        import cv2
    except Exception as e:  # pragma: no cover
        raise RuntimeError(f"no video decoder available for {path}: {e}")

    def count_frames() -> int:
        cap = cv2.VideoCapture(path)
        n = 0
        while cap.grab():
            n += 1
        cap.release()
        return n

    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise RuntimeError(f"could not open video {path}")
    src_fps = float(cap.get(cv2.CAP_PROP_FPS) or 0.0)
    n_src = int(cap.get(cv2.CAP_PROP_FRAME_COUNT) or 0)
    if n_src <= 0:                       # container without a frame count: one extra demux pass
        n_src = count_frames()
    if n_src <= 0:
        cap.release()
        raise RuntimeError(f"could not decode any frame from {path}")
    wanted = resample_indices(n_src, src_fps, float(fps)) if (fps is not None and src_fps > 0) else list(range(n_src))

    def convert(fr):
        h, w = fr.shape[:2]
        if resolution is not None and (h != resolution or w != resolution):
            sw, sh, x0, y0 = letterbox_geometry(w, h, resolution)
            small = cv2.resize(fr, (sw, sh), interpolation=cv2.INTER_CUBIC)
            canvas = np.zeros((resolution, resolution, 3), dtype=small.dtype)
            canvas[y0:y0 + sh, x0:x0 + sw] = small
            fr = canvas
        return torch.from_numpy(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB)).permute(2, 0, 1).contiguous()

    out, k, i, last = [], 0, 0, None
    while k < len(wanted):
        if wanted[k] > i:                # skip without decoding to pixels
            if not cap.grab():
                break
            i += 1
            continue
        ok, fr = cap.read()              # source frame i is wanted (possibly several times: fps above the source rate)
        if not ok:
            break
        last = convert(fr)
        while k < len(wanted) and wanted[k] == i:
            out.append(last)
            k += 1
        i += 1
    cap.release()
    if last is None:
        raise RuntimeError(f"could not decode any frame from {path}")
    while k < len(wanted):               # the container over-reported its length: hold the last frame (ffmpeg -r pads the same way)
        out.append(last)
        k += 1
    return torch.stack(out).contiguous()
