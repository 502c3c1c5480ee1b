"""CPU oracle for the VideoLLM-online per-frame hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product package never does (its hot path fails loudly without the CUDA library).

What it is: a plain-PyTorch, CPU restatement of the reference's algorithm for this path — the glue in
/root/reference (models/vision_live.py, models/modeling_live.py, models/live_llama/modeling_live_llama.py,
demo/inference.py) PLUS the arithmetic that lives in the un-vendored, un-pinned third-party dependency
HuggingFace `transformers` (5.5.0 in this image: models/siglip/modeling_siglip.py,
models/llama/modeling_llama.py, cache_utils.py, activations.py).  Each function cites the file:line it
follows (`HF:` = transformers).  Dtypes follow what the reference does on a CPU host: ViT in fp32
(`torch.cuda.amp.autocast` is a no-op without CUDA), decoder and connector in bf16.

Pinning: the reference ships NO golden vectors or numeric tests for this path, so parity
is pinned against outputs of the reference's own modules executed in the build container:
tests/golden/make_golden.py imports /root/reference + transformers, loads the same seeded weights and
writes tests/golden/*.pt; tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

import collections
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]


# =============================================================================== vision tower
def siglip_embeddings(vs: StateDict, pixel_values: torch.Tensor, patch: int) -> torch.Tensor:
    """SiglipVisionEmbeddings.forward, HF:models/siglip/modeling_siglip.py:175-186."""
    w, b = vs["embeddings.patch_embedding.weight"], vs["embeddings.patch_embedding.bias"]
    x = F.conv2d(pixel_values.to(w.dtype), w, b, stride=patch)           # [B, C, g, g]
    x = x.flatten(2).transpose(1, 2)                                     # [B, P, C]
    return x + vs["embeddings.position_embedding.weight"][None]


def _mha(x_q, x_kv, wq, bq, wk, bk, wv, bv, wo, bo, heads: int) -> torch.Tensor:
    """Multi-head attention, scale head_dim^-0.5, no mask (SiglipAttention.forward,
    HF:...siglip.py:275-312; nn.MultiheadAttention for the pooling head)."""
    B, Lq, C = x_q.shape
    Lk = x_kv.shape[1]
    hd = C // heads
    q = F.linear(x_q, wq, bq).view(B, Lq, heads, hd).transpose(1, 2)
    k = F.linear(x_kv, wk, bk).view(B, Lk, heads, hd).transpose(1, 2)
    v = F.linear(x_kv, wv, bv).view(B, Lk, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, scale=hd ** -0.5)
    o = o.transpose(1, 2).reshape(B, Lq, C)
    return F.linear(o, wo, bo)


def siglip_encoder_layer(vs: StateDict, i: int, h: torch.Tensor, heads: int, eps: float) -> torch.Tensor:
    """SiglipEncoderLayer.forward, HF:...siglip.py:340-362 (pre-LN, GELU-tanh MLP :323-327)."""
    p = f"encoder.layers.{i}."
    C = h.shape[-1]
    r = h
    x = F.layer_norm(h, (C,), vs[p + "layer_norm1.weight"], vs[p + "layer_norm1.bias"], eps)
    x = _mha(x, x, vs[p + "self_attn.q_proj.weight"], vs[p + "self_attn.q_proj.bias"],
             vs[p + "self_attn.k_proj.weight"], vs[p + "self_attn.k_proj.bias"],
             vs[p + "self_attn.v_proj.weight"], vs[p + "self_attn.v_proj.bias"],
             vs[p + "self_attn.out_proj.weight"], vs[p + "self_attn.out_proj.bias"], heads)
    h = r + x
    r = h
    x = F.layer_norm(h, (C,), vs[p + "layer_norm2.weight"], vs[p + "layer_norm2.bias"], eps)
    x = F.linear(x, vs[p + "mlp.fc1.weight"], vs[p + "mlp.fc1.bias"])
    x = F.gelu(x, approximate="tanh")
    x = F.linear(x, vs[p + "mlp.fc2.weight"], vs[p + "mlp.fc2.bias"])
    return r + x


def siglip_pool_head(vs: StateDict, last_hidden: torch.Tensor, heads: int, eps: float) -> torch.Tensor:
    """SiglipMultiheadAttentionPoolingHead.forward, HF:...siglip.py:639-651 -> pooler_output [B, C]."""
    B, _, C = last_hidden.shape
    probe = vs["head.probe"].repeat(B, 1, 1)
    ipw, ipb = vs["head.attention.in_proj_weight"], vs["head.attention.in_proj_bias"]
    x = _mha(probe, last_hidden, ipw[:C], ipb[:C], ipw[C:2 * C], ipb[C:2 * C], ipw[2 * C:], ipb[2 * C:],
             vs["head.attention.out_proj.weight"], vs["head.attention.out_proj.bias"], heads)
    r = x
    x = F.layer_norm(x, (C,), vs["head.layernorm.weight"], vs["head.layernorm.bias"], eps)
    x = F.linear(x, vs["head.mlp.fc1.weight"], vs["head.mlp.fc1.bias"])
    x = F.gelu(x, approximate="tanh")
    x = F.linear(x, vs["head.mlp.fc2.weight"], vs["head.mlp.fc2.bias"])
    return (r + x)[:, 0]


def siglip_vision_encode(vs: StateDict, cfg, frames_u8: torch.Tensor) -> torch.Tensor:
    """_siglip_vision_encode, models/vision_live.py:10-30: rescale + normalise, ViT, CLS := pooler_output,
    3x3 adaptive average pool of the patch-token grid, concat -> [B, frame_num_tokens, C] (fp32 on CPU)."""
    x = frames_u8 * 0.00392156862745098                                   # :12 (uint8 * float -> fp32)
    x = (x - 0.5) / 0.5                                                   # torchvision normalize, mean=std=.5
    heads, eps = cfg.vision_num_attention_heads, cfg.vision_layer_norm_eps
    h = siglip_embeddings(vs, x, cfg.vision_patch_size)
    for i in range(cfg.vision_num_hidden_layers):
        h = siglip_encoder_layer(vs, i, h, heads, eps)
    C = h.shape[-1]
    last = F.layer_norm(h, (C,), vs["post_layernorm.weight"], vs["post_layernorm.bias"], eps)   # HF:...siglip.py:618
    outs = []
    if cfg.frame_token_cls:
        outs.append(siglip_pool_head(vs, last, heads, eps)[:, None])       # models/vision_live.py:27
    if cfg.frame_token_pooled:
        s = int(math.sqrt(last.shape[1]))
        sp = F.adaptive_avg_pool2d(last.reshape(last.shape[0], s, s, C).permute(0, 3, 1, 2), tuple(cfg.frame_token_pooled))
        outs.append(sp.flatten(2, 3).permute(0, 2, 1))                     # :17-23
    return torch.cat(outs, dim=1)


def connector(sd: StateDict, tokens: torch.Tensor) -> torch.Tensor:
    """Linear -> GELUActivation(use_gelu_python) -> Linear in the model dtype
    (models/live_llama/modeling_live_llama.py:18-22; HF:activations.py:78-86: the truthy positional
    argument selects `x * 0.5 * (1 + erf(x / sqrt(2)))`, evaluated op by op in bf16)."""
    x = F.linear(tokens, sd["connector.0.weight"], sd["connector.0.bias"])
    x = x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    return F.linear(x, sd["connector.2.weight"], sd["connector.2.bias"])


def visual_embed(sd: StateDict, vs: Optional[StateDict], cfg, frames: torch.Tensor) -> torch.Tensor:
    """LiveMixin.visual_embed, models/modeling_live.py:21-27 -> [B * frame_num_tokens, hidden] bf16."""
    if vs is not None and frames.dtype == torch.uint8:
        frames = siglip_vision_encode(vs, cfg, frames)
    frames = frames.to(sd["connector.0.weight"].dtype)
    out = connector(sd, frames)
    return out.view(-1, out.shape[-1])


# =============================================================================== decoder
class KVCache:
    """DynamicCache restated (HF:cache_utils.py:88-121): per layer K, V [1, kv_heads, N, hd], grown by cat."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    def get_seq_length(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[-2]

    def __bool__(self):
        return self.get_seq_length() > 0

    def update(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=-2)
            self.v[layer] = torch.cat([self.v[layer], v], dim=-2)
        return self.k[layer], self.v[layer]

    def crop(self, n: int):
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i], self.v[i] = self.k[i][..., :n, :], self.v[i][..., :n, :]


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm.forward, HF:models/llama/modeling_llama.py:62-67."""
    dt = x.dtype
    x = x.to(torch.float32)
    x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return w * x.to(dt)


def rope_cos_sin(cfg, position_ids: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding.forward (default rope), HF:...llama.py:124-135."""
    d = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(dtype=torch.float) / d))
    freqs = (inv_freq[None, :, None].float() @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


class LoraAdapter:
    """An UNMERGED PEFT LoRA adapter, as the reference runs it at inference (models/modeling_live.py:203-216:
    PeftModel.from_pretrained(..., is_trainable=False), never merged).  peft is an un-vendored, un-pinned dependency
    (README.md:54); its published Linear forward (peft/tuners/lora/layer.py, `Linear.forward`) is
        result = base_layer(x);  result = result + lora_B(lora_A(dropout(x))) * scaling
    with dropout the identity in eval mode and every op in the module dtype (bf16).  `state` uses the adapter's
    own key names (base_model.model.<module>.lora_{A,B}[.<adapter>].weight)."""

    def __init__(self, state: StateDict, scaling: float):
        self.scaling = scaling
        self.ab = {}
        for k, a in state.items():
            if ".lora_A" in k:
                mod = k.split(".lora_A")[0]
                mod = mod[len("base_model.model."):] if mod.startswith("base_model.model.") else mod
                self.ab[mod] = (a, state[k.replace("lora_A", "lora_B")])

    def linear(self, x: torch.Tensor, sd: StateDict, module: str) -> torch.Tensor:
        y = F.linear(x, sd[module + ".weight"])
        if module in self.ab:
            a, b = self.ab[module]
            y = y + F.linear(F.linear(x, a), b) * self.scaling
        return y


class _NoLora:
    @staticmethod
    def linear(x, sd, module):
        return F.linear(x, sd[module + ".weight"])


def llama_forward(sd: StateDict, cfg, inputs_embeds: torch.Tensor, cache: KVCache, lora: Optional[LoraAdapter] = None,
                  return_hidden: bool = False) -> torch.Tensor:
    """KV-append forward of LlamaForCausalLM with a DynamicCache and sdpa attention
    (HF:...llama.py:375-426 model, :303-333 layer, :251-289 attention, :146-168 RoPE, :182-184 MLP,
    :485-487 lm_head on all positions).  inputs_embeds [q, H] (batch 1) -> logits [q, V].
    `lora`: optional unmerged adapter applied to every wrapped Linear (see LoraAdapter)."""
    lin = (lora or _NoLora).linear
    h = inputs_embeds[None]
    q_len = h.shape[1]
    past = cache.get_seq_length()
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    pos = torch.arange(past, past + q_len)[None]
    cos, sin = rope_cos_sin(cfg, pos, h.dtype)
    cos, sin = cos[:, None], sin[:, None]
    kv_len = past + q_len
    # causal mask with offset (HF:masking_utils.py:263-272): key j visible to query i iff j <= past + i
    mask = torch.arange(kv_len)[None, :] <= (past + torch.arange(q_len))[:, None]
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        r = h
        x = rms_norm(h, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = lin(x, sd, p + "self_attn.q_proj").view(1, q_len, nh, hd).transpose(1, 2)
        k = lin(x, sd, p + "self_attn.k_proj").view(1, q_len, nkv, hd).transpose(1, 2)
        v = lin(x, sd, p + "self_attn.v_proj").view(1, q_len, nkv, hd).transpose(1, 2)
        q = (q * cos) + (_rotate_half(q) * sin)
        k = (k * cos) + (_rotate_half(k) * sin)
        k_all, v_all = cache.update(i, k, v)
        g = nh // nkv
        kr = k_all[:, :, None].expand(1, nkv, g, kv_len, hd).reshape(1, nh, kv_len, hd)   # repeat_kv
        vr = v_all[:, :, None].expand(1, nkv, g, kv_len, hd).reshape(1, nh, kv_len, hd)
        a = F.scaled_dot_product_attention(q, kr, vr, attn_mask=mask[None, None], scale=hd ** -0.5)
        a = a.transpose(1, 2).reshape(1, q_len, nh * hd)
        h = r + lin(a, sd, p + "self_attn.o_proj")
        r = h
        x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        x = lin(F.silu(lin(x, sd, p + "mlp.gate_proj")) * lin(x, sd, p + "mlp.up_proj"), sd, p + "mlp.down_proj")
        h = r + x
    h = rms_norm(h, sd["model.norm.weight"], cfg.rms_norm_eps)
    if return_hidden:   # final-norm output [q, H] (test helper: crafting lm_head rows)
        return h[0]
    return lin(h, sd, "lm_head")[0]


def embed_tokens(sd: StateDict, ids: torch.Tensor) -> torch.Tensor:
    return F.embedding(ids, sd["model.embed_tokens.weight"])


def joint_embed(sd: StateDict, vs: Optional[StateDict], cfg, input_ids: Optional[torch.Tensor], frames: Optional[torch.Tensor]) -> torch.Tensor:
    """LiveMixin.joint_embed, models/modeling_live.py:29-42: embed `input_ids` (the <v> placeholder id lies one past
    the embedding table: clamped, then overwritten) and scatter visual_embed(frames) into the <v> positions."""
    if frames is None:
        return embed_tokens(sd, input_ids)
    if input_ids is None:
        return visual_embed(sd, vs, cfg, frames)
    emb = embed_tokens(sd, input_ids.clamp(max=cfg.vocab_size - 1)).clone()
    v_mask = input_ids == cfg.v_placeholder_id
    if v_mask.any():
        emb[v_mask] = visual_embed(sd, vs, cfg, frames)
    return emb


def decide(last_logits: torch.Tensor, interval_id: int, threshold: float) -> int:
    """demo/inference.py:76-79 on one logits row."""
    score = last_logits.view(1, 1, -1).softmax(dim=-1)
    if score[:, :, interval_id] < threshold:
        score[:, :, interval_id].zero_()
    return int(score.argmax(dim=-1))


def fast_greedy_generate(sd: StateDict, cfg, inputs_embeds: torch.Tensor, cache: KVCache, eos_token_id: int,
                         max_new: int = 100) -> List[int]:
    """models/modeling_live.py:173-182."""
    out = []
    for _ in range(max_new):
        logits = llama_forward(sd, cfg, inputs_embeds, cache)
        tok = int(logits[-1].argmax(dim=-1))
        out.append(tok)
        if tok == eos_token_id:
            break
        inputs_embeds = embed_tokens(sd, torch.tensor([tok]))
    return out


# =============================================================================== stream_evaluate
def stream_evaluate(sd: StateDict, vs: Optional[StateDict], cfg, input_ids: torch.Tensor, labels: torch.Tensor,
                    frames: torch.Tensor, ignore_token_id: int = -100, frame_token_interval_threshold: float = 0.0) -> torch.Tensor:
    """LiveMixin.stream_evaluate, models/modeling_live.py:44-168, batch 1: one teacher-forced forward over the whole
    conversation (all-position logits), then per turn LM perplexity / time difference / fluency / LM correctness; a turn
    whose frames are all predicted "silent" continues the stream on a cache trimmed to the turn's last frame token
    (:112-141; `trim_past_key_values(pkv, 0, stop)` == keep the first `stop` positions).
    Returns tensor [lm_ppl, frame_diff, fluency, lm_correctness]."""
    assert input_ids.size(0) == labels.size(0) == 1
    input_id, label = input_ids[0], labels[0]
    zero, one = torch.tensor(0, dtype=torch.int), torch.tensor(1, dtype=torch.int)
    turn_stops = ((input_id == cfg.eos_token_id).nonzero() + 1)[:, 0].tolist()
    turn_starts = [0] + turn_stops[:-1]
    num_turns = len(turn_starts)
    cache = KVCache(cfg.num_hidden_layers)
    logit = llama_forward(sd, cfg, joint_embed(sd, vs, cfg, input_id, frames), cache)
    v_id = cfg.v_placeholder_id
    use_interval = cfg.frame_token_interval_id is not None
    interval_id = cfg.frame_token_interval_id if use_interval else cfg.eos_token_id
    fnt = int(cfg.frame_token_cls) + (cfg.frame_token_pooled[0] * cfg.frame_token_pooled[1] if cfg.frame_token_pooled else 0)
    past_num_frames = 0
    lm_ppls, frame_diffs, fluencies, lm_correctness = [], [], [], []
    for r, (turn_start, turn_stop) in enumerate(zip(turn_starts, turn_stops)):
        turn_label = label[turn_start:turn_stop]
        turn_learn_mask = turn_label != ignore_token_id
        if not turn_learn_mask.any():
            continue
        turn_logit = logit[turn_start:turn_stop]
        turn_input_id = input_id[turn_start:turn_stop]
        turn_v_mask = turn_input_id == v_id
        turn_num_frames = turn_v_mask.sum() // fnt
        turn_stream_mask = turn_v_mask & turn_learn_mask
        turn_lm_mask = turn_learn_mask & ~turn_stream_mask
        if turn_lm_mask.any():                                                            # :86-96
            ml, mt = turn_logit[turn_lm_mask], turn_label[turn_lm_mask]
            lm_ppls.append(F.cross_entropy(ml, mt).exp())
            wrong = ml.argmax(dim=-1) != mt
            num_lm_correct_tokens = wrong.nonzero()[0, 0] if wrong.any() else (~wrong).sum()
            lm_correctness.append(num_lm_correct_tokens / mt.numel())
        if turn_stream_mask.any():                                                        # :99-142
            score = turn_logit.softmax(dim=-1)[turn_stream_mask]
            if frame_token_interval_threshold > 0:
                score[score[:, interval_id] < frame_token_interval_threshold] = 0
            pred = score.argmax(dim=-1) != interval_id
            if pred.any():
                frame_diff = turn_stream_mask.sum() - pred.nonzero()[0, 0] - 1
            else:
                last_stream_idx = turn_stream_mask.nonzero()[-1, 0]
                if r == num_turns - 1:
                    frame_diff = zero
                else:
                    next_nf = (input_id[turn_starts[r + 1]:turn_stops[r + 1]] == v_id).sum() // fnt
                    n_app = min(next_nf, turn_num_frames - 1)
                    if n_app == 0:
                        frame_diff = zero
                    else:
                        app_frames = frames[past_num_frames + turn_num_frames: past_num_frames + turn_num_frames + n_app]
                        ph = ([interval_id] if use_interval else []) + [v_id] * fnt
                        app_ids = torch.tensor(ph * int(n_app), dtype=torch.long)
                        trimmed = KVCache(cfg.num_hidden_layers)
                        stop = int(turn_start + last_stream_idx + 1)
                        for li in range(cfg.num_hidden_layers):
                            trimmed.k[li], trimmed.v[li] = cache.k[li][..., :stop, :], cache.v[li][..., :stop, :]
                        app_logit = llama_forward(sd, cfg, joint_embed(sd, vs, cfg, app_ids, app_frames), trimmed)
                        idxs = torch.arange(len(ph) - 1, len(app_ids), len(ph))
                        app_score = app_logit[idxs].softmax(dim=-1)
                        if frame_token_interval_threshold > 0:
                            app_score[app_score[:, interval_id] < frame_token_interval_threshold] = 0
                        app_pred = app_score.argmax(dim=-1) != interval_id
                        frame_diff = -(app_pred.nonzero()[0, 0] + 1) if app_pred.any() else -n_app
            frame_diffs.append(torch.as_tensor(frame_diff).abs())
        if turn_lm_mask.any() and turn_stream_mask.any():                                 # :145-154
            n_v = turn_stream_mask.sum()
            n_valid = mt.numel() + n_v
            if frame_diff == 0:
                fluency = (n_v + num_lm_correct_tokens) / n_valid
            elif frame_diff > 0:
                fluency = (n_v - frame_diff) / n_valid
            else:
                fluency = (n_v - 1) / n_valid
            fluencies.append(fluency)
        past_num_frames += turn_num_frames
    lm_ppl = torch.stack(lm_ppls).mean() if lm_ppls else one
    frame_diff = torch.stack(frame_diffs).float().mean() if frame_diffs else zero
    fluency = torch.stack([torch.as_tensor(f) for f in fluencies]).float().mean() if fluencies else one
    lm_c = torch.stack([torch.as_tensor(c) for c in lm_correctness]).float().mean() if lm_correctness else one
    return torch.stack([torch.as_tensor(lm_ppl).float(), torch.as_tensor(frame_diff).float(), torch.as_tensor(fluency).float(),
                        torch.as_tensor(lm_c).float()])


# =============================================================================== state machine
class OracleLiveInfer:
    """LiveInfer restated for CPU (demo/inference.py:12-124).  `logit_hook(logits_row, step_kind)` lets a
    test script the model's decisions (random weights never emit the "]\\n" / EOS protocol ids)."""

    def __init__(self, sd, vs, cfg, tokenizer, *, frame_fps: int, system_prompt: str, logit_hook=None):
        self.sd, self.vs, self.cfg, self.tokenizer = sd, vs, cfg, tokenizer
        self.frame_fps = frame_fps
        self.frame_num_tokens = cfg.frame_num_tokens
        self.frame_token_interval_id = cfg.frame_token_interval_id
        self.frame_token_interval_threshold = 0.725
        self.eos_token_id = cfg.eos_token_id
        self.max_new = 100
        self.logit_hook = logit_hook
        self._start_ids = tokenizer.apply_chat_template([{'role': 'system', 'content': system_prompt}], add_stream_prompt=True, return_tensors='pt')
        self._added_stream_prompt_ids = tokenizer.apply_chat_template([{}], add_stream_prompt=True, return_tensors='pt')
        self._added_stream_generation_ids = tokenizer.apply_chat_template([{}], add_stream_generation_prompt=True, return_tensors='pt')
        self.reset()

    def reset(self):
        self.query_queue, self.frame_embeds_queue = collections.deque(), collections.deque()
        self.video_time, self.last_frame_idx, self.video_tensor = 0, -1, None
        self.last_ids = torch.tensor([[]], dtype=torch.long)
        self.cache = KVCache(self.cfg.num_hidden_layers)
        self.trace = []

    def load_video(self, video_tensor):
        self.video_tensor = video_tensor
        self.num_video_frames = video_tensor.size(0)

    def input_query_stream(self, query, video_time=None):
        self.query_queue.append((self.video_time if video_time is None else video_time, query))

    def input_video_stream(self, video_time):
        frame_idx = int(video_time * self.frame_fps)
        if frame_idx > self.last_frame_idx:
            ranger = range(self.last_frame_idx + 1, frame_idx + 1)
            embeds = visual_embed(self.sd, self.vs, self.cfg, self.video_tensor[ranger.start:ranger.stop]).split(self.frame_num_tokens)
            self.frame_embeds_queue.extend([(r / self.frame_fps, e) for r, e in zip(ranger, embeds)])
        self.last_frame_idx, self.video_time = frame_idx, video_time

    def _logits(self, embeds, kind):
        logits = llama_forward(self.sd, self.cfg, embeds, self.cache)[-1]
        if self.logit_hook is not None:
            logits = self.logit_hook(logits, kind)
        return logits

    def _call_for_response(self, video_time, query):
        if query is not None:
            ids = self.tokenizer.apply_chat_template([{'role': 'user', 'content': query}], add_stream_query_prompt=True, add_generation_prompt=True, return_tensors='pt')
        else:
            assert int(self.last_ids) == self.cfg.stream_end_id
            ids = self._added_stream_generation_ids
        out = []
        embeds = embed_tokens(self.sd, ids[0])
        for _ in range(self.max_new):
            tok = int(self._logits(embeds, 'gen').argmax(dim=-1))
            out.append(tok)
            if tok == self.eos_token_id:
                break
            embeds = embed_tokens(self.sd, torch.tensor([tok]))
        self.last_ids = torch.tensor([[out[-1]]])
        self.trace.append(('response', video_time, list(out), self.cache.get_seq_length()))
        text = self.tokenizer.decode(out, skip_special_tokens=True)
        return (f'(Video Time = {video_time}s) User: {query}' if query else query), f'(Video Time = {video_time}s) Assistant:{text}'

    def _call_for_streaming(self):
        while self.frame_embeds_queue:
            if self.query_queue and self.frame_embeds_queue[0][0] > self.query_queue[0][0]:
                return self.query_queue.popleft()
            video_time, frame_embeds = self.frame_embeds_queue.popleft()
            if not self.cache:
                self.last_ids = self._start_ids
            elif self.last_ids.numel() == 1 and int(self.last_ids) == self.eos_token_id:
                self.last_ids = torch.cat([self.last_ids, self._added_stream_prompt_ids], dim=1)
            embeds = torch.cat([embed_tokens(self.sd, self.last_ids.view(-1)), frame_embeds.view(-1, frame_embeds.shape[-1])], 0)
            logits = self._logits(embeds, 'frame')
            if self.query_queue and video_time >= self.query_queue[0][0]:
                return self.query_queue.popleft()
            nxt = decide(logits, self.frame_token_interval_id, self.frame_token_interval_threshold)
            self.last_ids = torch.tensor([[nxt]])
            self.trace.append(('frame', video_time, nxt, self.cache.get_seq_length()))
            if nxt != self.frame_token_interval_id:
                return video_time, None
        return None, None

    def __call__(self):
        video_time, query = self._call_for_streaming()
        response = None
        if video_time is not None:
            query, response = self._call_for_response(video_time, query)
        return query, response
