"""Weights: synthetic HF-style state dicts, LoRA merge, and packing into the engine layout.

The reference loads `meta-llama/Meta-Llama-3-8B-Instruct` + a PEFT adapter (LoRA r=128, alpha=256 on
q,k,v,o,gate,up,down,lm_head plus the `connector` module) + SigLIP-L (models/modeling_live.py:200-220,
models/arguments_live.py:16-19).  No checkpoint is available offline, so tests and benchmarks use
seeded synthetic tensors with the reference's parameter names and shapes; the
same dict feeds the oracle, the HF reference modules (golden generation) and — after `pack_for_engine`
— the CUDA engine.

Engine layout (name -> tensor), all contiguous on the engine's device:
  embed [V,H] bf16 | final_norm [H] bf16 | lm_head [V,H] bf16 | rope.cos / rope.sin [max_pos,64] bf16
  L{i}.in_norm, L{i}.post_norm [H] bf16 | L{i}.qkv [(nh+2nkv)*128, H] bf16 (q|k|v rows)
  L{i}.o [H, nh*128] | L{i}.gate_up [2I, H] (tile-interleaved: 64 gate rows of features 64j.., then the matching 64 up
  rows, per 128-row tile j - one MMA tile then holds both operands of SwiGLU) | L{i}.down [H, I]   (bf16)
  conn.0.w [H,C] bf16, conn.0.b [H] f32, conn.2.w [H,H] bf16, conn.2.b [H] f32
  vit.patch.w [C, 3*ps*ps] f16, vit.patch.b [C] f32, vit.pos [P,C] f32, vit.post_ln.{w,b} f32
  vit.L{i}.{ln1,ln2}.{w,b} f32, vit.L{i}.qkv.w [3C,C] f16 (+ .b f32), out.w [C,C], fc1.w [M,C], fc2.w [C,M]
  vit.head.q [C] f32 (= in_proj_q(probe), fp16-rounded), vit.head.kv.{w,b}, out.{w,b}, ln.{w,b}, fc1, fc2
16-bit biases are stored as fp32 copies of their 16-bit values (the GEMM epilogue adds them in fp32).
"""
from __future__ import annotations

import math
import re
from typing import Dict

import torch

from .config import LiveConfig

StateDict = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- synthetic HF-style
def _gen(shape, std, gen, device, dtype, mean=0.0):
    t = torch.empty(shape, device=device, dtype=torch.float32)
    t.normal_(mean, std, generator=gen)
    return t.to(dtype)


def synthetic_llm_state(cfg: LiveConfig, seed: int = 0, device="cpu", std: float = 0.02) -> StateDict:
    """`LiveLlamaForCausalLM.state_dict()`-shaped random weights (bf16), LoRA already merged."""
    g = torch.Generator(device=device).manual_seed(seed)
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    nh, nkv, hd, C = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.vision_hidden_size
    bf = torch.bfloat16
    sd: StateDict = {}
    sd["model.embed_tokens.weight"] = _gen((V, H), 1.0, g, device, bf)
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        sd[p + "self_attn.q_proj.weight"] = _gen((nh * hd, H), std, g, device, bf)
        sd[p + "self_attn.k_proj.weight"] = _gen((nkv * hd, H), std, g, device, bf)
        sd[p + "self_attn.v_proj.weight"] = _gen((nkv * hd, H), std, g, device, bf)
        sd[p + "self_attn.o_proj.weight"] = _gen((H, nh * hd), std, g, device, bf)
        sd[p + "mlp.gate_proj.weight"] = _gen((I, H), std, g, device, bf)
        sd[p + "mlp.up_proj.weight"] = _gen((I, H), std, g, device, bf)
        sd[p + "mlp.down_proj.weight"] = _gen((H, I), std, g, device, bf)
        sd[p + "input_layernorm.weight"] = _gen((H,), 0.1, g, device, bf, mean=1.0)
        sd[p + "post_attention_layernorm.weight"] = _gen((H,), 0.1, g, device, bf, mean=1.0)
    sd["model.norm.weight"] = _gen((H,), 0.1, g, device, bf, mean=1.0)
    sd["lm_head.weight"] = _gen((V, H), std, g, device, bf)
    sd["connector.0.weight"] = _gen((H, C), std, g, device, bf)
    sd["connector.0.bias"] = _gen((H,), std, g, device, bf)
    sd["connector.2.weight"] = _gen((H, H), std, g, device, bf)
    sd["connector.2.bias"] = _gen((H,), std, g, device, bf)
    return sd


def synthetic_vision_state(cfg: LiveConfig, seed: int = 1, device="cpu", std: float = 0.02) -> StateDict:
    """`SiglipVisionModel(...).vision_model.state_dict()`-shaped random weights (fp32, as loaded by
    build_live_vision, models/vision_live.py:55)."""
    g = torch.Generator(device=device).manual_seed(seed)
    C, M, ps = cfg.vision_hidden_size, cfg.vision_intermediate_size, cfg.vision_patch_size
    P = cfg.num_patches
    f32 = torch.float32
    sd: StateDict = {}
    sd["embeddings.patch_embedding.weight"] = _gen((C, 3, ps, ps), std, g, device, f32)
    sd["embeddings.patch_embedding.bias"] = _gen((C,), std, g, device, f32)
    sd["embeddings.position_embedding.weight"] = _gen((P, C), std, g, device, f32)
    for i in range(cfg.vision_num_hidden_layers):
        p = f"encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            sd[p + ln + ".weight"] = _gen((C,), 0.1, g, device, f32, mean=1.0)
            sd[p + ln + ".bias"] = _gen((C,), std, g, device, f32)
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = _gen((C, C), std, g, device, f32)
            sd[p + f"self_attn.{nm}.bias"] = _gen((C,), std, g, device, f32)
        sd[p + "mlp.fc1.weight"] = _gen((M, C), std, g, device, f32)
        sd[p + "mlp.fc1.bias"] = _gen((M,), std, g, device, f32)
        sd[p + "mlp.fc2.weight"] = _gen((C, M), std, g, device, f32)
        sd[p + "mlp.fc2.bias"] = _gen((C,), std, g, device, f32)
    sd["post_layernorm.weight"] = _gen((C,), 0.1, g, device, f32, mean=1.0)
    sd["post_layernorm.bias"] = _gen((C,), std, g, device, f32)
    sd["head.probe"] = _gen((1, 1, C), 1.0, g, device, f32)
    sd["head.attention.in_proj_weight"] = _gen((3 * C, C), std, g, device, f32)
    sd["head.attention.in_proj_bias"] = _gen((3 * C,), std, g, device, f32)
    sd["head.attention.out_proj.weight"] = _gen((C, C), std, g, device, f32)
    sd["head.attention.out_proj.bias"] = _gen((C,), std, g, device, f32)
    sd["head.layernorm.weight"] = _gen((C,), 0.1, g, device, f32, mean=1.0)
    sd["head.layernorm.bias"] = _gen((C,), std, g, device, f32)
    sd["head.mlp.fc1.weight"] = _gen((M, C), std, g, device, f32)
    sd["head.mlp.fc1.bias"] = _gen((M,), std, g, device, f32)
    sd["head.mlp.fc2.weight"] = _gen((C, M), std, g, device, f32)
    sd["head.mlp.fc2.bias"] = _gen((C,), std, g, device, f32)
    return sd


# --------------------------------------------------------------------------- LoRA merge (K20)
_LORA_RE = re.compile(r"^(?:base_model\.model\.)?(.*)\.lora_A(?:\.[^.]+)?\.weight$")


def merge_lora(llm_state: StateDict, adapter_state: StateDict, lora_alpha: float = 256, lora_r: int = 128) -> StateDict:
    """W' = W + (alpha/r) * B @ A for every wrapped Linear, plus `modules_to_save` (connector) overrides.

    The reference keeps the adapter unmerged at inference (models/modeling_live.py:216,
    y = W x + 2.0 * B(A(x))); merging at load removes 2 small GEMMs per Linear from the hot path.
    The product is formed in fp32 and rounded once to bf16 (parity note in DESIGN.md)."""
    out = dict(llm_state)
    scaling = float(lora_alpha) / float(lora_r)
    for key, a in adapter_state.items():
        m = _LORA_RE.match(key)
        if m:
            base = m.group(1) + ".weight"
            bkey = key.replace("lora_A", "lora_B")
            if base not in out or bkey not in adapter_state:
                raise KeyError(f"LoRA tensor {key} has no base weight {base} / partner {bkey}")
            w = out[base]
            delta = adapter_state[bkey].float() @ a.float()
            out[base] = (w.float() + scaling * delta.to(w.device)).to(w.dtype)
        elif "modules_to_save" in key or re.match(r"^(?:base_model\.model\.)?connector\.", key):
            name = re.sub(r"^(?:base_model\.model\.)?", "", key)
            name = re.sub(r"\.modules_to_save\.[^.]+", "", name)
            if name.startswith("connector."):
                out[name] = adapter_state[key].to(torch.bfloat16)
    return out


# --------------------------------------------------------------------------- engine layout
def rope_tables(cfg: LiveConfig, n_positions: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    """bf16 cos/sin tables exactly as LlamaRotaryEmbedding.forward produces them for
    position_ids = arange(n) (HF:models/llama/modeling_llama.py:124-135): fp32 outer product,
    fp32 cos/sin, cast to the activation dtype.  Only the first head_dim/2 columns are stored
    (HF concatenates two identical halves)."""
    d = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(dtype=torch.float) / d))
    pos = torch.arange(n_positions, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    return (freqs.cos().to(torch.bfloat16).contiguous().to(device),
            freqs.sin().to(torch.bfloat16).contiguous().to(device))


def _b32_from16(t: torch.Tensor, dt16) -> torch.Tensor:
    return t.to(dt16).to(torch.float32).contiguous()


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I, H] + [I, H] -> [2I, H] with rows ordered per 128-row tile: gate[64j:64j+64] then up[64j:64j+64]."""
    I, H = gate.shape
    if I % 64 != 0:
        raise ValueError("intermediate_size must be a multiple of 64")
    return torch.stack([gate.view(I // 64, 64, H), up.view(I // 64, 64, H)], 1).reshape(2 * I, H).contiguous()


def pack_llm_for_engine(cfg: LiveConfig, sd: StateDict, device, max_positions: int) -> StateDict:
    bf = torch.bfloat16
    out: StateDict = {}

    def take(k):
        return sd[k].to(device=device, dtype=bf).contiguous()

    out["embed"] = take("model.embed_tokens.weight")
    out["final_norm"] = take("model.norm.weight")
    out["lm_head"] = take("lm_head.weight")
    cos, sin = rope_tables(cfg, max_positions, device)
    out["rope.cos"], out["rope.sin"] = cos, sin
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        out[f"L{i}.in_norm"] = take(p + "input_layernorm.weight")
        out[f"L{i}.post_norm"] = take(p + "post_attention_layernorm.weight")
        out[f"L{i}.qkv"] = torch.cat([take(p + "self_attn.q_proj.weight"), take(p + "self_attn.k_proj.weight"),
                                      take(p + "self_attn.v_proj.weight")], 0).contiguous()
        out[f"L{i}.o"] = take(p + "self_attn.o_proj.weight")
        out[f"L{i}.gate_up"] = interleave_gate_up(take(p + "mlp.gate_proj.weight"), take(p + "mlp.up_proj.weight"))
        out[f"L{i}.down"] = take(p + "mlp.down_proj.weight")
    if "connector.0.weight" in sd:
        out["conn.0.w"] = take("connector.0.weight")
        out["conn.0.b"] = _b32_from16(sd["connector.0.bias"].to(device), bf)
        out["conn.2.w"] = take("connector.2.weight")
        out["conn.2.b"] = _b32_from16(sd["connector.2.bias"].to(device), bf)
    return out


def pack_vision_for_engine(cfg: LiveConfig, sd: StateDict, device) -> StateDict:
    f16, f32 = torch.float16, torch.float32
    C = cfg.vision_hidden_size
    out: StateDict = {}

    def w16(k):
        return sd[k].to(device=device, dtype=f16).contiguous()

    def b32(k):  # bias as the fp16 autocast copy would hold it
        return _b32_from16(sd[k].to(device), f16)

    def f(k):
        return sd[k].to(device=device, dtype=f32).contiguous()

    out["vit.patch.w"] = w16("embeddings.patch_embedding.weight").reshape(C, -1).contiguous()
    out["vit.patch.b"] = b32("embeddings.patch_embedding.bias")
    out["vit.pos"] = f("embeddings.position_embedding.weight")
    out["vit.post_ln.w"], out["vit.post_ln.b"] = f("post_layernorm.weight"), f("post_layernorm.bias")
    for i in range(cfg.vision_num_hidden_layers):
        p, q = f"encoder.layers.{i}.", f"vit.L{i}."
        out[q + "ln1.w"], out[q + "ln1.b"] = f(p + "layer_norm1.weight"), f(p + "layer_norm1.bias")
        out[q + "ln2.w"], out[q + "ln2.b"] = f(p + "layer_norm2.weight"), f(p + "layer_norm2.bias")
        out[q + "qkv.w"] = torch.cat([w16(p + f"self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous()
        out[q + "qkv.b"] = torch.cat([b32(p + f"self_attn.{n}_proj.bias") for n in "qkv"], 0).contiguous()
        out[q + "out.w"], out[q + "out.b"] = w16(p + "self_attn.out_proj.weight"), b32(p + "self_attn.out_proj.bias")
        out[q + "fc1.w"], out[q + "fc1.b"] = w16(p + "mlp.fc1.weight"), b32(p + "mlp.fc1.bias")
        out[q + "fc2.w"], out[q + "fc2.b"] = w16(p + "mlp.fc2.weight"), b32(p + "mlp.fc2.bias")
    if cfg.frame_token_cls:
        ipw, ipb = sd["head.attention.in_proj_weight"].to(device), sd["head.attention.in_proj_bias"].to(device)
        # q of the constant probe, as the fp16 autocast Linear would produce it (fp32 accumulate, one rounding)
        probe16 = sd["head.probe"].to(device).reshape(1, C).to(f16).float()
        q = probe16 @ ipw[:C].to(f16).float().t() + ipb[:C].to(f16).float()
        out["vit.head.q"] = q.reshape(C).to(f16).to(f32).contiguous()
        out["vit.head.kv.w"] = ipw[C:].to(f16).contiguous()
        out["vit.head.kv.b"] = _b32_from16(ipb[C:], f16)
        out["vit.head.out.w"], out["vit.head.out.b"] = w16("head.attention.out_proj.weight"), b32("head.attention.out_proj.bias")
        out["vit.head.ln.w"], out["vit.head.ln.b"] = f("head.layernorm.weight"), f("head.layernorm.bias")
        out["vit.head.fc1.w"], out["vit.head.fc1.b"] = w16("head.mlp.fc1.weight"), b32("head.mlp.fc1.bias")
        out["vit.head.fc2.w"], out["vit.head.fc2.b"] = w16("head.mlp.fc2.weight"), b32("head.mlp.fc2.bias")
    return out


def synthetic_engine_weights(cfg: LiveConfig, device, max_positions: int, seed: int = 0, std: float = 0.02) -> StateDict:
    """Full-size synthetic weights generated directly in the engine layout on `device` (bench path: no
    16 GB host round trip).  Same distributions as synthetic_*_state, different random stream."""
    g = torch.Generator(device=device).manual_seed(seed)
    bf, f16, f32 = torch.bfloat16, torch.float16, torch.float32
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    C, M, ps, P = cfg.vision_hidden_size, cfg.vision_intermediate_size, cfg.vision_patch_size, cfg.num_patches
    out: StateDict = {}
    out["embed"] = _gen((V, H), 1.0, g, device, bf)
    out["final_norm"] = _gen((H,), 0.1, g, device, bf, 1.0)
    out["lm_head"] = _gen((V, H), std, g, device, bf)
    out["rope.cos"], out["rope.sin"] = rope_tables(cfg, max_positions, device)
    for i in range(cfg.num_hidden_layers):
        out[f"L{i}.in_norm"] = _gen((H,), 0.1, g, device, bf, 1.0)
        out[f"L{i}.post_norm"] = _gen((H,), 0.1, g, device, bf, 1.0)
        out[f"L{i}.qkv"] = _gen(((nh + 2 * nkv) * hd, H), std, g, device, bf)
        out[f"L{i}.o"] = _gen((H, nh * hd), std, g, device, bf)
        out[f"L{i}.gate_up"] = _gen((2 * I, H), std, g, device, bf)
        out[f"L{i}.down"] = _gen((H, I), std, g, device, bf)
    out["conn.0.w"] = _gen((H, C), std, g, device, bf)
    out["conn.0.b"] = _gen((H,), std, g, device, bf).float()
    out["conn.2.w"] = _gen((H, H), std, g, device, bf)
    out["conn.2.b"] = _gen((H,), std, g, device, bf).float()
    out["vit.patch.w"] = _gen((C, 3 * ps * ps), std, g, device, f16)
    out["vit.patch.b"] = _gen((C,), std, g, device, f16).float()
    out["vit.pos"] = _gen((P, C), std, g, device, f32)
    out["vit.post_ln.w"], out["vit.post_ln.b"] = _gen((C,), 0.1, g, device, f32, 1.0), _gen((C,), std, g, device, f32)
    for i in range(cfg.vision_num_hidden_layers):
        q = f"vit.L{i}."
        for ln in ("ln1", "ln2"):
            out[q + ln + ".w"], out[q + ln + ".b"] = _gen((C,), 0.1, g, device, f32, 1.0), _gen((C,), std, g, device, f32)
        out[q + "qkv.w"], out[q + "qkv.b"] = _gen((3 * C, C), std, g, device, f16), _gen((3 * C,), std, g, device, f16).float()
        out[q + "out.w"], out[q + "out.b"] = _gen((C, C), std, g, device, f16), _gen((C,), std, g, device, f16).float()
        out[q + "fc1.w"], out[q + "fc1.b"] = _gen((M, C), std, g, device, f16), _gen((M,), std, g, device, f16).float()
        out[q + "fc2.w"], out[q + "fc2.b"] = _gen((C, M), std, g, device, f16), _gen((C,), std, g, device, f16).float()
    if cfg.frame_token_cls:
        out["vit.head.q"] = _gen((C,), 0.5, g, device, f16).float()
        out["vit.head.kv.w"], out["vit.head.kv.b"] = _gen((2 * C, C), std, g, device, f16), _gen((2 * C,), std, g, device, f16).float()
        out["vit.head.out.w"], out["vit.head.out.b"] = _gen((C, C), std, g, device, f16), _gen((C,), std, g, device, f16).float()
        out["vit.head.ln.w"], out["vit.head.ln.b"] = _gen((C,), 0.1, g, device, f32, 1.0), _gen((C,), std, g, device, f32)
        out["vit.head.fc1.w"], out["vit.head.fc1.b"] = _gen((M, C), std, g, device, f16), _gen((M,), std, g, device, f16).float()
        out["vit.head.fc2.w"], out["vit.head.fc2.b"] = _gen((C, M), std, g, device, f16), _gen((C,), std, g, device, f16).float()
    return out


def engine_weight_bytes(weights: StateDict) -> int:
    return sum(t.numel() * t.element_size() for t in weights.values())
