"""Python handle on one `vlo_engine` (one per GPU / process).

Thin: torch owns the weight storage (so torch.distributed/NCCL can broadcast it) and the
caller-visible tensors; every compute call goes through the C ABI in include/vlo_b200.h.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import VloConfig, VloDecision, VloError, check
from .config import LiveConfig

DECISION_DTYPE_FIELDS = 8  # 8 x 4 bytes


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Decision:
    """Host copy of one `vlo_decision`."""
    __slots__ = ("argmax_id", "argmax_excl_id", "p_interval", "max_logit", "top2_margin", "lse", "argmax_prob_id")

    def __init__(self, row_i32: torch.Tensor, row_f32: torch.Tensor):
        self.argmax_id = int(row_i32[0])
        self.argmax_excl_id = int(row_i32[1])
        self.p_interval = float(row_f32[2])
        self.max_logit = float(row_f32[3])
        self.top2_margin = float(row_f32[4])
        self.lse = float(row_f32[5])
        self.argmax_prob_id = int(row_i32[6])

    def next_id(self, interval_id: int, threshold: float) -> int:
        """The reference's rule (demo/inference.py:76-79) on bf16 probabilities: if p(interval) < threshold the
        interval id is zeroed before the argmax.  torch compares the bf16 tensor with the Python float
        after casting the scalar to bf16, so do the same."""
        thr = float(torch.tensor(threshold, dtype=torch.bfloat16))
        if self.p_interval < thr:
            return self.argmax_excl_id
        return self.argmax_prob_id


class Engine:
    def __init__(self, cfg: LiveConfig, device: str | torch.device = "cuda:0", *, max_streams: int = 1,
                 max_kv_tokens: int = 16384, max_step_tokens: int = 128, max_vit_batch: int = 8):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise VloError("the vlo_b200 engine runs on a CUDA (sm_100a) device only; there is no CPU path")
        if not torch.cuda.is_available():
            raise VloError("no CUDA device visible: the hot path cannot run (no fallback by design)")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        c = VloConfig()
        c.hidden_size, c.num_layers = cfg.hidden_size, cfg.num_hidden_layers
        c.num_heads, c.num_kv_heads, c.head_dim = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        c.intermediate_size, c.vocab_size, c.rms_norm_eps = cfg.intermediate_size, cfg.vocab_size, cfg.rms_norm_eps
        c.vit_hidden, c.vit_layers = cfg.vision_hidden_size, cfg.vision_num_hidden_layers
        c.vit_heads, c.vit_mlp = cfg.vision_num_attention_heads, cfg.vision_intermediate_size
        c.image_size, c.patch_size, c.vit_ln_eps = cfg.frame_resolution, cfg.vision_patch_size, cfg.vision_layer_norm_eps
        c.frame_token_cls = 1 if cfg.frame_token_cls else 0
        c.pool_h, c.pool_w = (cfg.frame_token_pooled or [0, 0])
        c.max_streams, c.max_kv_tokens = max_streams, max_kv_tokens
        c.max_step_tokens, c.max_vit_batch = max_step_tokens, max_vit_batch
        self.max_streams, self.max_kv_tokens = max_streams, max_kv_tokens
        self.max_step_tokens, self.max_vit_batch = max_step_tokens, max_vit_batch
        self._cfg_struct = c
        self._h = C.c_void_p()
        self._lock = threading.Lock()
        self._weights: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(self.index):
            check(self.lib.vlo_engine_create(C.byref(c), self.index, C.byref(self._h)), "vlo_engine_create")
        # device scratch for results (decisions, last logits)
        self._dec_dev = torch.zeros(max_streams, DECISION_DTYPE_FIELDS, dtype=torch.int32, device=self.device)
        self._dec_host = torch.zeros(max_streams, DECISION_DTYPE_FIELDS, dtype=torch.int32).pin_memory()
        self._logits = torch.zeros(max_streams, cfg.vocab_size, dtype=torch.bfloat16, device=self.device)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.vlo_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_weights(self, weights: Dict[str, torch.Tensor]):
        """Register tensors already in the engine layout (weights.pack_*_for_engine)."""
        for name, t in weights.items():
            if t.device != self.device:
                t = t.to(self.device)
            t = t.contiguous()
            self._weights[name] = t  # keep the storage alive
            check(self.lib.vlo_load_tensor(self._h, name.encode(), _ptr(t), t.numel() * t.element_size()),
                  f"vlo_load_tensor({name})")
        check(self.lib.vlo_finalize_weights(self._h), "vlo_finalize_weights")

    @property
    def weights(self) -> Dict[str, torch.Tensor]:
        return self._weights

    def device_bytes(self) -> int:
        return int(self.lib.vlo_engine_device_bytes(self._h))

    # ------------------------------------------------------------------ streams
    def stream_open(self) -> int:
        sid = C.c_int(-1)
        check(self.lib.vlo_stream_open(self._h, C.byref(sid)), "vlo_stream_open")
        return sid.value

    def stream_reset(self, sid: int):
        check(self.lib.vlo_stream_reset(self._h, sid), "vlo_stream_reset")

    def stream_close(self, sid: int):
        check(self.lib.vlo_stream_close(self._h, sid), "vlo_stream_close")

    def kv_len(self, sid: int) -> int:
        n = C.c_int(0)
        check(self.lib.vlo_kv_len(self._h, sid, C.byref(n)), "vlo_kv_len")
        return n.value

    def kv_truncate(self, sid: int, new_len: int):
        check(self.lib.vlo_kv_truncate(self._h, sid, new_len), "vlo_kv_truncate")

    def kv_copy_prefix(self, src_sid: int, dst_sid: int, n_tokens: int):
        check(self.lib.vlo_kv_copy_prefix(self._h, src_sid, dst_sid, n_tokens, self._stream()), "vlo_kv_copy_prefix")

    def kv_fill_synthetic(self, sid: int, n_tokens: int, seed: int = 0):
        check(self.lib.vlo_kv_fill_synthetic(self._h, sid, n_tokens, seed, self._stream()), "vlo_kv_fill_synthetic")

    def kv_read(self, sid: int, layer: int, is_v: bool) -> torch.Tensor:
        n = self.kv_len(sid)
        out = torch.empty(self.cfg.num_key_value_heads, n, self.cfg.head_dim, dtype=torch.bfloat16, device=self.device)
        check(self.lib.vlo_kv_read(self._h, sid, layer, int(is_v), _ptr(out), self._stream()), "vlo_kv_read")
        return out

    def kv_write(self, sid: int, layer: int, is_v: bool, rows: torch.Tensor):
        rows = rows.to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert rows.shape[0] == self.cfg.num_key_value_heads and rows.shape[2] == self.cfg.head_dim
        check(self.lib.vlo_kv_write(self._h, sid, layer, int(is_v), _ptr(rows), rows.shape[1], self._stream()), "vlo_kv_write")

    # ------------------------------------------------------------------ hot path
    def vit_encode(self, frames_u8: torch.Tensor, *, return_vit_tokens: bool = False, connector: bool = True):
        """uint8 [B,3,S,S] -> bf16 [B*frame_num_tokens, hidden] (visual_embed, models/modeling_live.py:21-27)."""
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
            raise VloError("vit_encode expects uint8 frames [B,3,S,S]")
        # pinned host frames (a live feed) are copied asynchronously on the caller's stream
        frames_u8 = frames_u8.to(self.device, non_blocking=frames_u8.device.type == "cpu" and frames_u8.is_pinned()).contiguous()
        B = frames_u8.shape[0]
        S = self.cfg.frame_resolution
        if tuple(frames_u8.shape[1:]) != (3, S, S):
            raise VloError(f"frames must be [B,3,{S},{S}], got {tuple(frames_u8.shape)}")
        nt = self.cfg.frame_num_tokens
        outs, toks = [], []
        for b0 in range(0, B, self.max_vit_batch):
            fb = frames_u8[b0:b0 + self.max_vit_batch]
            n = fb.shape[0]
            out = torch.empty(n * nt, self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device) if connector else None
            vt = torch.empty(n, nt, self.cfg.vision_hidden_size, dtype=torch.float32, device=self.device) if return_vit_tokens else None
            check(self.lib.vlo_vit_encode(self._h, _ptr(fb), n, _ptr(out), _ptr(vt), self._stream()), "vlo_vit_encode")
            outs.append(out)
            toks.append(vt)
        out = torch.cat(outs, 0) if connector else None
        if return_vit_tokens:
            return out, torch.cat(toks, 0)
        return out

    def connector(self, tokens: torch.Tensor) -> torch.Tensor:
        tokens = tokens.to(device=self.device, dtype=torch.bfloat16).contiguous().view(-1, self.cfg.vision_hidden_size)
        out = torch.empty(tokens.shape[0], self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
        check(self.lib.vlo_connector(self._h, _ptr(tokens), tokens.shape[0], _ptr(out), self._stream()), "vlo_connector")
        return out

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        out = torch.empty(ids.numel(), self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
        if ids.numel():
            check(self.lib.vlo_embed_tokens(self._h, _ptr(ids), ids.numel(), _ptr(out), self._stream()), "vlo_embed_tokens")
        return out

    def step(self, stream_ids: Sequence[int], q_lens: Sequence[int], embeds: torch.Tensor, *,
             row_ids: Optional[torch.Tensor] = None, interval_id: Optional[int] = None,
             want_logits: bool = True):
        """KV-append forward of a ragged batch.  `row_ids` (int64 [sum q_lens], optional): id >= 0 -> the row is the
        embedding of that token (gathered on the device), id < 0 -> the row of `embeds` is used as is.  Returns (logits [n_seqs, V] bf16 view or None, decisions
        device tensor int32 [n_seqs, 8]).  Nothing is synchronised; call `read_decisions` for host values."""
        n = len(stream_ids)
        T = int(sum(q_lens))
        embeds = embeds.contiguous()
        if T > self.max_step_tokens:
            return self._step_chunked(stream_ids, q_lens, embeds, row_ids, interval_id, want_logits)
        if embeds.dtype != torch.bfloat16 or embeds.device != self.device or embeds.numel() != T * self.cfg.hidden_size:
            raise VloError(f"step: embeds must be bf16 [{T},{self.cfg.hidden_size}] on {self.device}")
        sid = (C.c_int32 * n)(*stream_ids)
        ql = (C.c_int32 * n)(*q_lens)
        iid = self.cfg.frame_token_interval_id if interval_id is None else interval_id
        if iid is None:
            iid = -1
        logits = self._logits[:n] if want_logits else None
        pid = None
        if row_ids is not None:
            pid = row_ids.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
            if pid.numel() != T:
                raise VloError(f"step: row_ids must have one entry per packed row ({T}), got {pid.numel()}")
        check(self.lib.vlo_step_ids(self._h, n, sid, ql, _ptr(pid), _ptr(embeds), _ptr(logits), _ptr(self._dec_dev), iid,
                                    self._stream()), "vlo_step")
        return logits, self._dec_dev[:n]

    def _step_chunked(self, stream_ids, q_lens, embeds, row_ids, interval_id, want_logits):
        """Inputs longer than max_step_tokens (the first frame's system prompt, long queries): feed the KV-append
        forward in pieces — chunked streaming is exactly one causal pass.  Each
        sequence's logits / decision come from the sub-step that holds its last piece."""
        n, cap, H = len(stream_ids), self.max_step_tokens, self.cfg.hidden_size
        if row_ids is not None:  # materialise the gathered token rows once; the sub-steps then see plain embeddings
            pid = row_ids.to(self.device).view(-1)
            rows = self.embed_tokens(pid.clamp(min=0))
            embeds = torch.where((pid >= 0)[:, None], rows, embeds)
        logits_out = torch.empty(n, self.cfg.vocab_size, dtype=torch.bfloat16, device=self.device) if want_logits else None
        dec_out = torch.empty(n, DECISION_DTYPE_FIELDS, dtype=torch.int32, device=self.device)
        off, done = 0, [0] * n
        starts = [int(sum(q_lens[:i])) for i in range(n)]
        while any(done[i] < q_lens[i] for i in range(n)):
            sub_ids, sub_lens, sub_rows, room = [], [], [], cap
            for i in range(n):
                left = q_lens[i] - done[i]
                if left <= 0 or room <= 0:
                    continue
                take = min(left, room)
                sub_ids.append(i)
                sub_lens.append(take)
                sub_rows.append(embeds[starts[i] + done[i]: starts[i] + done[i] + take])
                room -= take
            lg, dc = self.step([stream_ids[i] for i in sub_ids], sub_lens, torch.cat(sub_rows, 0), interval_id=interval_id,
                               want_logits=want_logits)
            for j, i in enumerate(sub_ids):
                done[i] += sub_lens[j]
                if done[i] == q_lens[i]:
                    dec_out[i].copy_(dc[j])
                    if want_logits:
                        logits_out[i].copy_(lg[j])
        self._dec_dev[:n].copy_(dec_out)
        if want_logits:
            self._logits[:n].copy_(logits_out)
        return (self._logits[:n] if want_logits else None), self._dec_dev[:n]

    def read_decisions(self, n: int) -> List[Decision]:
        """One small D2H copy + sync: the only host<->device sync of a frame step."""
        self._dec_host[:n].copy_(self._dec_dev[:n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        f32 = self._dec_host.view(torch.float32)
        return [Decision(self._dec_host[i], f32[i]) for i in range(n)]

    def last_step_logits(self, n_tokens: int) -> torch.Tensor:
        out = torch.empty(n_tokens, self.cfg.vocab_size, dtype=torch.bfloat16, device=self.device)
        check(self.lib.vlo_last_step_logits(self._h, _ptr(out), self._stream()), "vlo_last_step_logits")
        return out

    def last_step_hidden(self, n_tokens: int) -> torch.Tensor:
        out = torch.empty(n_tokens, self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
        check(self.lib.vlo_last_step_hidden(self._h, _ptr(out), self._stream()), "vlo_last_step_hidden")
        return out

    def launch_count(self) -> int:
        return int(self.lib.vlo_launch_count())
