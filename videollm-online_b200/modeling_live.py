"""Host-side mirror of the reference's model API for the inference path.

Keeps the names `demo/inference.py` calls — `build_model_and_tokenizer`, `fast_greedy_generate`,
`model.config`, `model.to`, `model.get_input_embeddings()`, `model.visual_embed`, `model.joint_embed`,
`model(inputs_embeds=..., past_key_values=..., use_cache=True)` — over the CUDA engine
(reference: models/modeling_live.py:11-42,173-222; models/live_llama/modeling_live_llama.py:11-73).
All arithmetic happens in libvlo_b200.so; this file only moves handles around.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .config import LiveArguments, LiveConfig, llama3_8b_siglip_l
from .engine import Engine, VloError
from .tokenization_live import build_live_tokenizer_and_update_config
from . import weights as W


class StreamKV:
    """The `past_key_values` object: an opaque handle on one engine-owned KV stream.  Falsy until it
    holds tokens, like `None` / an empty DynamicCache (demo/inference.py:61,98)."""

    def __init__(self, engine: Engine, stream_id: int):
        self.engine, self.stream_id = engine, stream_id

    def get_seq_length(self) -> int:
        return self.engine.kv_len(self.stream_id)

    def __len__(self):
        return self.get_seq_length()

    def __bool__(self):
        return self.get_seq_length() > 0

    def crop(self, max_length: int):
        """trim_past_key_values(0, max_length) (models/modeling_live.py:170-171)."""
        self.engine.kv_truncate(self.stream_id, max_length)
        return self


@dataclass
class LiveOutput:
    """CausalLMOutputWithPast look-alike.  `logits` holds the LAST position only ([B, 1, V]); the
    reference computes all positions but reads only [:, -1:] (demo/inference.py:76,
    models/modeling_live.py:177).  `decisions` is the device-side summary of that row."""
    logits: torch.Tensor
    past_key_values: StreamKV
    decisions: torch.Tensor
    loss: Optional[torch.Tensor] = None


class _Embedding:
    def __init__(self, engine: Engine):
        self.engine = engine

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        shape = tuple(ids.shape)
        return self.engine.embed_tokens(ids).view(*shape, self.engine.cfg.hidden_size)

    @property
    def weight(self):
        return self.engine.weights["embed"]


class LiveLlamaForCausalLM:
    """Engine-backed stand-in for models.live_llama.LiveLlamaForCausalLM (inference surface only)."""

    def __init__(self, config: LiveConfig, engine: Engine):
        self.config = config
        self.engine = engine
        self.dtype = torch.bfloat16
        self.vocab_size = config.vocab_size
        self._embed = _Embedding(engine)
        self._default_stream: Optional[int] = None
        self.has_vision = any(k.startswith("vit.") for k in engine.weights)

    # --- torch-module idioms the reference uses
    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device) != self.engine.device and torch.device(device).type == "cuda" \
                and torch.device(device).index not in (None, self.engine.index):
            raise VloError("an engine is pinned to its GPU; build another engine for another device")
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    @property
    def device(self):
        return self.engine.device

    def get_input_embeddings(self):
        return self._embed

    # --- LiveMixin
    def set_vision_inside(self):
        if not self.has_vision:
            raise VloError("vision tower weights were not loaded")

    def visual_embed(self, frames: torch.Tensor) -> torch.Tensor:
        """models/modeling_live.py:21-27.  uint8 frames -> ViT + connector; float features
        ([N, frame_num_tokens, vision_hidden], pre-extracted) -> connector only."""
        if frames.dtype == torch.uint8:
            if not self.has_vision:
                raise VloError("visual_embed got raw frames but no vision tower is loaded")
            return self.engine.vit_encode(frames)
        return self.engine.connector(frames)

    def joint_embed(self, input_ids: torch.Tensor = None, frames: torch.Tensor = None) -> torch.Tensor:
        """models/modeling_live.py:29-42."""
        if frames is None:
            return self.get_input_embeddings()(input_ids)
        if input_ids is None:
            return self.visual_embed(frames)
        embeds = self.get_input_embeddings()(input_ids.clamp(max=self.vocab_size - 1))
        v_mask = input_ids.to(embeds.device) == self.config.v_placeholder_id
        if v_mask.any():
            embeds[v_mask] = self.visual_embed(frames)
        return embeds

    def new_stream(self) -> StreamKV:
        return StreamKV(self.engine, self.engine.stream_open())

    def __call__(self, input_ids: torch.Tensor = None, frames: torch.Tensor = None, inputs_embeds: torch.Tensor = None,
                 past_key_values: Optional[StreamKV] = None, use_cache: bool = True, return_dict: bool = True, **_):
        """KV-append forward (models/live_llama/modeling_live_llama.py:24-67), batch 1."""
        if inputs_embeds is None:
            inputs_embeds = self.joint_embed(input_ids, frames)
        if inputs_embeds.dim() == 3:
            if inputs_embeds.shape[0] != 1:
                raise VloError("model(...) is the reference's batch-1 API; use engine.step for ragged batches")
            inputs_embeds = inputs_embeds[0]
        if past_key_values is None:
            if self._default_stream is None:
                self._default_stream = self.engine.stream_open()
            else:
                self.engine.stream_reset(self._default_stream)
            past_key_values = StreamKV(self.engine, self._default_stream)
        q = inputs_embeds.shape[0]
        logits, dec = self.engine.step([past_key_values.stream_id], [q], inputs_embeds.to(torch.bfloat16))
        return LiveOutput(logits=logits.view(1, 1, -1), past_key_values=past_key_values, decisions=dec)

    forward = __call__


def fast_greedy_generate(*, model: LiveLlamaForCausalLM, inputs_embeds: torch.Tensor, past_key_values: StreamKV,
                         eos_token_id: int, inplace_output_ids: torch.Tensor):
    """models/modeling_live.py:173-182: greedy AR until EOS or the buffer is full.  The argmax is taken
    on the device; one 32-byte read-back per token replaces the reference's tensor `if` sync."""
    eng = model.engine
    n = 0
    for i in range(inplace_output_ids.size(1)):
        out = model(inputs_embeds=inputs_embeds, past_key_values=past_key_values, use_cache=True)
        past_key_values = out.past_key_values
        new_id = eng.read_decisions(1)[0].argmax_id
        inplace_output_ids[:, i] = new_id
        n = i + 1
        if new_id == eos_token_id:
            break
        inputs_embeds = model.get_input_embeddings()(torch.tensor([[new_id]], device=eng.device))
    return inplace_output_ids[:, :n], past_key_values


def build_live(*, is_training: bool = False, config: Optional[LiveConfig] = None, llm_pretrained: str = None,
               resume_from_checkpoint: str = "", set_vision_inside: bool = False, device: str = "cuda:0",
               synthetic_weights: bool = False, seed: int = 0, max_streams: int = 1, max_kv_tokens: int = 16384,
               max_step_tokens: int = 128, max_vit_batch: int = 8, llm_state=None, vision_state=None,
               lora_r: int = 128, lora_alpha: int = 256, **kwargs):
    """models/modeling_live.py:184-222 for inference: build config + tokenizer, load (or synthesise)
    weights, merge LoRA, hand everything to the engine."""
    if is_training:
        raise NotImplementedError("the B200 engine covers the streaming-inference path only")
    cfg = config or llama3_8b_siglip_l()
    for k in ("frame_resolution", "frame_token_cls", "frame_token_pooled", "frame_num_tokens", "frame_token_interval",
              "vision_pretrained"):
        if kwargs.get(k) is not None and hasattr(cfg, k):
            setattr(cfg, k, kwargs[k])
    tokenizer = build_live_tokenizer_and_update_config(llm_pretrained or "", cfg)
    engine = Engine(cfg, device, max_streams=max_streams, max_kv_tokens=max_kv_tokens,
                    max_step_tokens=max_step_tokens, max_vit_batch=max_vit_batch)
    if llm_state is not None:
        packed = W.pack_llm_for_engine(cfg, llm_state, engine.device, max_kv_tokens)
        if set_vision_inside:
            if vision_state is None:
                raise VloError("set_vision_inside=True needs vision_state")
            packed.update(W.pack_vision_for_engine(cfg, vision_state, engine.device))
    elif synthetic_weights:
        packed = W.synthetic_engine_weights(cfg, engine.device, max_kv_tokens, seed=seed)
        if not set_vision_inside:
            packed = {k: v for k, v in packed.items() if not k.startswith("vit.")}
    else:
        packed = _load_checkpoints(cfg, llm_pretrained, resume_from_checkpoint, set_vision_inside, engine.device,
                                   max_kv_tokens, lora_r, lora_alpha)
    engine.load_weights(packed)
    model = LiveLlamaForCausalLM(cfg, engine)
    return model, tokenizer


def _load_checkpoints(cfg, llm_pretrained, adapter, with_vision, device, max_positions, lora_r, lora_alpha):
    """Real weights: base Llama safetensors + PEFT adapter (+ SigLIP), local files only."""
    import glob
    import os
    try:
        from safetensors.torch import load_file
    except Exception as e:  # pragma: no cover
        raise VloError(f"safetensors unavailable: {e}")
    if not llm_pretrained or not os.path.isdir(llm_pretrained):
        raise VloError(f"checkpoint directory '{llm_pretrained}' not found (no network here); pass "
                       "--synthetic_weights true to run with seeded random weights")
    sd = {}
    for f in sorted(glob.glob(os.path.join(llm_pretrained, "*.safetensors"))):
        sd.update(load_file(f))
    if adapter:
        ad = {}
        for f in sorted(glob.glob(os.path.join(adapter, "*.safetensors"))):
            ad.update(load_file(f))
        sd = W.merge_lora(sd, ad, lora_alpha=lora_alpha, lora_r=lora_r)
    packed = W.pack_llm_for_engine(cfg, sd, device, max_positions)
    if with_vision:
        vdir = cfg.vision_pretrained
        vs = {}
        for f in sorted(glob.glob(os.path.join(vdir, "*.safetensors"))):
            vs.update(load_file(f))
        vs = {k[len("vision_model."):]: v for k, v in vs.items() if k.startswith("vision_model.")}
        packed.update(W.pack_vision_for_engine(cfg, vs, device))
    return packed


def build_live_llama(**kwargs):
    """models/live_llama/modeling_live_llama.py:72-73."""
    return build_live(**kwargs)


build_model_and_tokenizer = build_live_llama
