"""Streaming session state machine — the engine-backed `LiveInfer`.

Same public surface and protocol as the reference's demo/inference.py:12-124 (`load_video`,
`input_video_stream`, `input_query_stream`, `__call__`, `reset`, `frame_token_interval_threshold`) so
demo/cli.py and demo/app.py can drive it unchanged.  Differences, all behind the same results:
  * the device is a parameter, not a hard-coded 'cuda';
  * the speak/silent decision and the greedy argmax are computed on the device; one 32-byte read-back
    per frame / generated token replaces the reference's two tensor-`if` syncs (demo/inference.py:77,80);
  * the steady-state frame step feeds [interval-token id | 10 frame embeddings] and lets the engine gather
    the token row, so no separate embedding launch or torch.cat is needed.
"""
from __future__ import annotations

import collections
import logging
import threading
from dataclasses import asdict

import torch

from .config import LiveArguments, parse_args
from .modeling_live import LiveLlamaForCausalLM, StreamKV, build_model_and_tokenizer, fast_greedy_generate

logger = logging.getLogger("liveinfer")


class LiveInfer:
    def __init__(self, args: LiveArguments = None, *, model: LiveLlamaForCausalLM = None, tokenizer=None,
                 device: str = None, stream: StreamKV = None) -> None:
        args = args or parse_args()
        if model is None:
            kw = asdict(args)
            kw["device"] = device or args.device
            model, tokenizer = build_model_and_tokenizer(is_training=False, set_vision_inside=True, **kw)
        self.model, self.tokenizer = model, tokenizer
        self.device = model.device
        cfg = model.config
        # visual
        self.hidden_size = cfg.hidden_size
        self.frame_fps = args.frame_fps
        self.frame_interval = 1 / self.frame_fps
        self.frame_resolution = cfg.frame_resolution
        self.frame_num_tokens = cfg.frame_num_tokens
        self.frame_v_placeholder = cfg.v_placeholder * self.frame_num_tokens
        self.frame_token_interval_id = cfg.frame_token_interval_id
        self.frame_placeholder_ids = torch.tensor(cfg.v_placeholder_id).repeat(cfg.frame_num_tokens).reshape(1, -1)
        # generation
        self.system_prompt = args.system_prompt
        self.inplace_output_ids = torch.zeros(1, 100, dtype=torch.long)  # host buffer: ids are decided on device, read back per token
        self.frame_token_interval_threshold = 0.725
        self.eos_token_id = cfg.eos_token_id
        self.stream_end_id = cfg.stream_end_id
        tok = self.tokenizer
        self._start_ids = tok.apply_chat_template([{'role': 'system', 'content': self.system_prompt}], add_stream_prompt=True, return_tensors='pt')
        self._added_stream_prompt_ids = tok.apply_chat_template([{}], add_stream_prompt=True, return_tensors='pt')
        self._added_stream_generation_ids = tok.apply_chat_template([{}], add_stream_generation_prompt=True, return_tensors='pt')
        self._kv = stream if stream is not None else model.new_stream()
        # test seam: decision_hook(decision, call_index) -> decision lets a test script the model's choices
        # (random weights never emit the "]\\n" / EOS protocol ids); None in production.
        self.decision_hook = None
        self._n_calls = 0
        # encode-ahead: with a loaded clip the ViT + connector of the NEXT frame run on a side CUDA stream while the
        # decoder works on the current one (same per-frame arithmetic; a live stream delivers the next frame
        # during the current step anyway).  Set prefetch_next = False for strictly sequential behaviour.
        self.prefetch_next = True
        # frames encoded per encode-ahead call.  1 (default): the next frame only, through the small-tile ViT configuration
        # that shares the SMs with the decoder step (measured 211 frames/s in the pipelined loop, versus 202 for groups of 4).
        # With a loaded clip a larger depth sends the next `prefetch_depth` frames through ONE batched ViT pass (the
        # reference itself batches every frame between two input_video_stream calls, demo/inference.py:106); from 3 frames
        # on that pass uses the 2-CTA tensor-bound GEMMs, which the engine serialises against decoder steps.
        self.prefetch_depth = 1
        self._side = torch.cuda.Stream(self.device)
        # demo/app.py drives input_video_stream and __call__ from separate Gradio callbacks: both touch the
        # engine's single set of ViT workspaces and the encode-ahead state, so their bodies are serialised.
        self._lock = threading.RLock()
        self._alloc_step_buffers(256)
        self._row_ids_cache = {}
        self._prefetched = {}       # frame_idx -> (embeds [frame_num_tokens, hidden], event)
        self._prefetch_event = None   # last event recorded on the side stream (it owns the engine's ViT workspaces)
        self._want_prefetch = None  # first frame index to encode ahead right after the next decoder step is enqueued
        self.reset()

    def _alloc_step_buffers(self, rows: int):
        self._packed = torch.zeros(rows, self.hidden_size, dtype=torch.bfloat16, device=self.device)
        self._row_ids_host = torch.zeros(rows, dtype=torch.int64).pin_memory()
        self._row_ids_dev = torch.zeros(rows, dtype=torch.int64, device=self.device)

    def _drop_prefetch(self):
        """Forget encoded-ahead frames; an encode-ahead still in flight owns the ViT workspaces: order behind it."""
        if getattr(self, "_prefetch_event", None) is not None:
            torch.cuda.current_stream(self.device).wait_event(self._prefetch_event)
        self._prefetched = {}
        self._prefetch_event = None
        self._want_prefetch = None

    # ------------------------------------------------------------------ session control
    def reset(self):
        self.query_queue = collections.deque()
        self.frame_embeds_queue = collections.deque()
        self.video_time = 0
        self.last_frame_idx = -1
        self.video_tensor = None
        self.last_ids = torch.tensor([[]], dtype=torch.long)
        self._kv.engine.stream_reset(self._kv.stream_id)
        self.past_key_values = None
        self._drop_prefetch()

    def load_video(self, video_path_or_tensor, keep_on_host: bool = False):
        """Reference: read_video(...)[0].to('cuda') (demo/inference.py:111-115).  Accepts a uint8
        [T,3,H,W] tensor directly (synthetic clips) or a path decodable by torchvision/cv2.
        keep_on_host=True leaves the clip in pinned host memory (a live feed): every frame is then copied to the GPU
        asynchronously by the step that encodes it instead of the whole clip being resident."""
        if isinstance(video_path_or_tensor, torch.Tensor):
            vt = video_path_or_tensor
        else:
            vt = _read_video_tchw(video_path_or_tensor, fps=self.frame_fps, resolution=self.frame_resolution)
        with self._lock:
            # an encode-ahead of the previous clip must neither be consumed nor keep reading the old tensor
            self._drop_prefetch()
            if keep_on_host:
                self.video_tensor = vt if (vt.device.type == "cpu" and vt.is_pinned()) else vt.cpu().pin_memory()
            else:
                self.video_tensor = vt.to(self.device)
        self.num_video_frames = self.video_tensor.size(0)
        self.video_duration = self.video_tensor.size(0) / self.frame_fps
        logger.warning(f'{"tensor" if isinstance(video_path_or_tensor, torch.Tensor) else video_path_or_tensor} -> '
                       f'{tuple(self.video_tensor.shape)}, {self.frame_fps} FPS')

    def input_query_stream(self, query, history=None, video_time=None):
        self.query_queue.append((self.video_time if video_time is None else video_time, query))
        if not self.past_key_values:
            return f'(NOTE: No video stream here. Please select or upload a video. Then the assistant will answer "{query} (at {self.video_time}s)" in the video stream)'
        return f'(NOTE: Received "{query}" (at {self.video_time}s). Please wait until previous frames have been processed)'

    def input_video_stream(self, video_time):
        with self._lock:
            self._input_video_stream(video_time)

    def _input_video_stream(self, video_time):
        frame_idx = int(video_time * self.frame_fps)
        if frame_idx > self.last_frame_idx:
            ranger = range(self.last_frame_idx + 1, frame_idx + 1)
            main = torch.cuda.current_stream(self.device)
            embeds, r = [], ranger.start
            while r < ranger.stop:
                hit = self._prefetched.pop(r, None)
                if hit is not None:                     # encoded ahead on the side stream
                    pe, ev = hit
                    main.wait_event(ev)
                    pe.record_stream(main)
                    embeds.append(pe)
                    r += 1
                    continue
                r2 = r                                  # contiguous run of frames that were not encoded ahead
                while r2 < ranger.stop and r2 not in self._prefetched:
                    r2 += 1
                if self._prefetch_event is not None:    # a side-stream ViT may still own the engine's ViT workspaces
                    main.wait_event(self._prefetch_event)
                embeds.extend(self.model.visual_embed(self.video_tensor[r:r2]).split(self.frame_num_tokens))
                r = r2
            for k in [k for k in self._prefetched if k <= frame_idx]:   # skipped-over frames
                del self._prefetched[k]
            self.frame_embeds_queue.extend([(r / self.frame_fps, e) for r, e in zip(ranger, embeds)])
            nxt = frame_idx + 1
            self._want_prefetch = nxt if (self.prefetch_next and self.video_tensor is not None and nxt not in self._prefetched
                                          and nxt < self.video_tensor.size(0)) else None
        self.last_frame_idx = frame_idx
        self.video_time = video_time

    # ------------------------------------------------------------------ decoder steps
    def _forward(self, ids: torch.Tensor, frame_embeds: torch.Tensor = None):
        """One KV-append step over [embed(ids) ; frame_embeds]; returns the device decision (host copy)."""
        with self._lock:
            return self._forward_locked(ids, frame_embeds)

    def _forward_locked(self, ids: torch.Tensor, frame_embeds: torch.Tensor = None):
        eng = self.model.engine
        id_list = [int(x) for x in ids.reshape(-1).tolist()]
        n_ids = len(id_list)
        n_fr = 0 if frame_embeds is None else frame_embeds.shape[0]
        T = n_ids + n_fr
        # pre-allocated step buffers: no per-step torch.empty / torch.cat / pageable H2D on the hot path
        if T > self._packed.shape[0]:
            self._alloc_step_buffers(2 * T)
        packed = self._packed[:T]
        if n_fr:
            packed[n_ids:] = frame_embeds.view(-1, self.hidden_size)
        key = (tuple(id_list), n_fr)
        row_ids = self._row_ids_cache.get(key) if n_fr else None   # steady state: [interval id | 10 frame rows]
        if row_ids is None:
            host = self._row_ids_host[:T]
            host[:n_ids] = torch.tensor(id_list, dtype=torch.int64)
            host[n_ids:] = -1
            if n_fr:    # frame steps repeat (same prefix ids): keep their device copy
                row_ids = host.to(self.device, non_blocking=True)
                if len(self._row_ids_cache) < 64:
                    self._row_ids_cache[key] = row_ids
            else:       # prompts / AR tokens: one pinned -> device copy, consumed before the decision read-back below
                row_ids = self._row_ids_dev[:T]
                row_ids.copy_(host, non_blocking=True)
        main = torch.cuda.current_stream(self.device)
        pre = None
        if self._want_prefetch is not None:
            pre = torch.cuda.Event()
            pre.record(main)
        eng.step([self._kv.stream_id], [T], packed, row_ids=row_ids)   # token rows gathered on the device
        self.past_key_values = self._kv
        if pre is not None:
            # encode-ahead of the next frame, enqueued AFTER the step's launches (the host is the critical path right
            # after a decision read-back) but ordered only behind what preceded the step: it runs concurrently with it
            nxt, self._want_prefetch = self._want_prefetch, None
            stop = min(nxt + max(1, int(self.prefetch_depth)), self.video_tensor.size(0))
            self._side.wait_event(pre)
            with torch.cuda.stream(self._side):
                pe = self.model.visual_embed(self.video_tensor[nxt:stop]).split(self.frame_num_tokens)
                ev = torch.cuda.Event()
                ev.record(self._side)
            for k, e in zip(range(nxt, stop), pe):
                self._prefetched[k] = (e, ev)
            self._prefetch_event = ev
        dec = eng.read_decisions(1)[0]
        if self.decision_hook is not None:
            dec = self.decision_hook(dec, self._n_calls)
        self._n_calls += 1
        return dec

    def _call_for_response(self, video_time, query):
        if query is not None:
            self.last_ids = self.tokenizer.apply_chat_template([{'role': 'user', 'content': query}], add_stream_query_prompt=True, add_generation_prompt=True, return_tensors='pt')
        else:
            assert int(self.last_ids) == self.stream_end_id, f'{self.last_ids} != {self.stream_end_id}'  # "]\n" closes the frame list
            self.last_ids = self._added_stream_generation_ids
        ids = self.last_ids
        n = 0
        for i in range(self.inplace_output_ids.size(1)):  # fast_greedy_generate, models/modeling_live.py:173-182
            new_id = self._forward(ids).argmax_id
            self.inplace_output_ids[0, i] = new_id
            n = i + 1
            if new_id == self.eos_token_id:
                break
            ids = torch.tensor([[new_id]], dtype=torch.long)
        output_ids = self.inplace_output_ids[:, :n]
        self.last_ids = output_ids[:, -1:].clone()
        if query:
            query = f'(Video Time = {video_time}s) User: {query}'
        response = f'(Video Time = {video_time}s) Assistant:{self.tokenizer.decode(output_ids[0], skip_special_tokens=True, clean_up_tokenization_spaces=True)}'
        return query, response

    def _call_for_streaming(self):
        while self.frame_embeds_queue:
            # 1. a query that is due before the next frame is answered first
            if self.query_queue and self.frame_embeds_queue[0][0] > self.query_queue[0][0]:
                video_time, query = self.query_queue.popleft()
                return video_time, query
            video_time, frame_embeds = self.frame_embeds_queue.popleft()
            if not self.past_key_values:
                self.last_ids = self._start_ids
            elif int(self.last_ids.reshape(-1)[-1]) == self.eos_token_id and self.last_ids.numel() == 1:
                self.last_ids = torch.cat([self.last_ids.reshape(1, -1), self._added_stream_prompt_ids], dim=1)
            dec = self._forward(self.last_ids, frame_embeds)
            # 2. a query due at this frame's time is answered right after the frame
            if self.query_queue and video_time >= self.query_queue[0][0]:
                video_time, query = self.query_queue.popleft()
                return video_time, query
            # 3. speak/silent: below-threshold interval probability -> the argmax excludes the interval id
            next_id = dec.next_id(self.frame_token_interval_id, self.frame_token_interval_threshold)
            self.last_ids = torch.tensor([[next_id]], dtype=torch.long)
            if next_id != self.frame_token_interval_id:
                return video_time, None
        return None, None

    def __call__(self):
        while not self.frame_embeds_queue:
            continue
        video_time, query = self._call_for_streaming()
        response = None
        if video_time is not None:
            query, response = self._call_for_response(video_time, query)
        return query, response


def _read_video_tchw(path: str, fps=None, resolution=None) -> torch.Tensor:
    """demo/inference.py:111-115 + the preprocessing demo/cli.py:15-20 does before it (see video_ingest.py)."""
    from .video_ingest import read_video_resampled
    return read_video_resampled(path, fps=fps, resolution=resolution)
