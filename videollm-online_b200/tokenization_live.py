"""Tokenizer side of the streaming protocol.

`render_chat` restates the reference's Jinja chat template (models/tokenization_live.py:27-65) in plain
Python so the prompt pieces `LiveInfer` needs ("\\n[", "]\\nAssistant:", "]\\nUser: ...") can be produced
without the Llama-3 tokenizer files, which are not available offline.  With a real tokenizer directory
`build_live_tokenizer_and_update_config` behaves like the reference's (adds `<v>`, derives the ids).
`ByteTokenizer` is a deterministic stand-in used by tests / benchmarks with synthetic weights.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from .config import LiveConfig


def render_chat(messages: List[dict], *, bos_token: str, eos_token: str, stream_placeholder=None,
                add_generation_prompt=False, add_stream_prompt=False, add_stream_generation_prompt=False,
                add_stream_query_prompt=False) -> str:
    """Text the reference template renders (models/tokenization_live.py:38-62).  Note that the
    template indexes messages[0]['role'], so an empty dict message (`[{}]`, demo/inference.py:34-35)
    renders only the trailing prompt."""
    out = []
    msgs = list(messages)
    if msgs and msgs[0].get("role") == "system":
        out.append(bos_token + msgs[0]["content"] + "\n")
        msgs = msgs[1:]
    for m in msgs:
        role = m.get("role")
        if role == "user":
            out.append(("]\nUser: " if add_stream_query_prompt else "\nUser: ") + m["content"])
        elif role == "assistant":
            out.append("\nAssistant: " + m["content"] + eos_token)
        elif role == "stream" and m.get("num_frames", 0) > 0:
            ph = stream_placeholder(m["num_frames"]) if stream_placeholder else ""
            out.append("\n[" + ph + "]")
    if add_generation_prompt:
        out.append("\nAssistant:")
    elif add_stream_prompt:
        out.append("\n[")
    elif add_stream_generation_prompt:
        out.append("]\nAssistant:")
    return "".join(out)


class ByteTokenizer:
    """Deterministic byte-level tokenizer with the special ids the protocol needs.

    ids: bytes map to 16 + byte (so they never collide with the specials below); the strings the
    Llama-3 tokenizer merges into single tokens on this path are single ids here too:
      "," -> frame_token_interval_id, "]\\n" -> stream_end_id, BOS / EOS -> config ids.
    """

    def __init__(self, cfg: LiveConfig):
        self.cfg = cfg
        self.bos_token, self.eos_token = "<|bos|>", "<|eos|>"
        self.bos_token_id, self.eos_token_id = cfg.bos_token_id, cfg.eos_token_id
        self._special = {self.bos_token: cfg.bos_token_id, self.eos_token: cfg.eos_token_id,
                         "]\n": cfg.stream_end_id, cfg.v_placeholder: cfg.v_placeholder_id}
        if cfg.frame_token_interval:
            self._special[cfg.frame_token_interval] = cfg.frame_token_interval_id
        self._inv = {v: k for k, v in self._special.items()}
        self._byte_base = 16
        reserved = set(self._special.values())
        # byte b -> id; skip ids taken by specials
        self._b2i, nxt = {}, self._byte_base
        for b in range(256):
            while nxt in reserved:
                nxt += 1
            self._b2i[b] = nxt
            nxt += 1
        if nxt > cfg.vocab_size:
            raise ValueError("vocab too small for the byte tokenizer")
        self._i2b = {v: k for k, v in self._b2i.items()}

    def __len__(self):
        return self.cfg.vocab_size + 1  # + <v>

    def encode(self, text: str) -> List[int]:
        ids, i = [], 0
        specials = sorted(self._special, key=len, reverse=True)
        while i < len(text):
            for s in specials:
                if text.startswith(s, i):
                    ids.append(self._special[s])
                    i += len(s)
                    break
            else:
                for b in text[i].encode("utf-8"):
                    ids.append(self._b2i[b])
                i += 1
        return ids

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True, clean_up_tokenization_spaces: bool = True) -> str:
        buf, out = bytearray(), []
        for t in (int(x) for x in ids):
            if t in self._i2b:
                buf.append(self._i2b[t])
                continue
            if buf:
                out.append(buf.decode("utf-8", errors="replace"))
                buf = bytearray()
            if t in self._inv:
                tok = self._inv[t]
                if not (skip_special_tokens and tok in (self.bos_token, self.eos_token, self.cfg.v_placeholder)):
                    out.append(tok)
        if buf:
            out.append(buf.decode("utf-8", errors="replace"))
        return "".join(out)

    def convert_tokens_to_ids(self, tok: str) -> int:
        return self._special.get(tok, self.encode(tok)[0])

    def apply_chat_template(self, messages, add_generation_prompt=False, add_stream_prompt=False,
                            add_stream_generation_prompt=False, add_stream_query_prompt=False,
                            return_tensors: Optional[str] = None, tokenize: bool = True, **_):
        ph = lambda n: self.cfg.frame_token_interval.join([self.cfg.frame_num_tokens * self.cfg.v_placeholder] * n)
        text = render_chat(messages, bos_token=self.bos_token, eos_token=self.eos_token, stream_placeholder=ph,
                           add_generation_prompt=add_generation_prompt, add_stream_prompt=add_stream_prompt,
                           add_stream_generation_prompt=add_stream_generation_prompt,
                           add_stream_query_prompt=add_stream_query_prompt)
        if not tokenize:
            return text
        ids = self.encode(text)
        if return_tensors == "pt":
            return torch.tensor([ids], dtype=torch.long)
        return ids


def build_live_tokenizer_and_update_config(llm_pretrained: str, model_config: LiveConfig):
    """models/tokenization_live.py:110-122.  Uses the HF tokenizer when `llm_pretrained` names a local checkpoint
    directory.  The byte tokenizer (ids then come from the config, never hard-coded) is used ONLY when no checkpoint is
    named (synthetic weights): with real weights a missing / unreadable tokenizer must fail loudly, as the reference
    does, instead of feeding byte ids to a trained embedding table."""
    import os
    if not llm_pretrained or not os.path.isdir(llm_pretrained):
        return ByteTokenizer(model_config)
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(llm_pretrained, use_fast=True, padding_side="left", local_files_only=True)
    tok.add_special_tokens({"additional_special_tokens": [model_config.v_placeholder]})
    model_config.v_placeholder_id = len(tok) - 1
    model_config.frame_token_interval_id = (tok.convert_tokens_to_ids(model_config.frame_token_interval)
                                            if model_config.frame_token_interval else None)
    model_config.eos_token_id = tok.eos_token_id
    end_ids = tok.encode("]\n", add_special_tokens=False)
    if len(end_ids) == 1:
        model_config.stream_end_id = end_ids[0]
    tok.pad_token = tok.eos_token
    ph = f"'{model_config.frame_token_interval}'.join([{model_config.frame_num_tokens} * '{model_config.v_placeholder}'] * message['num_frames'])"
    tok.chat_template = _jinja_template(ph)
    return tok


def _jinja_template(stream_placeholder_jinja2: str) -> str:
    # same rendering rules as render_chat, for HF tokenizers' apply_chat_template
    return (
        "{% if messages[0]['role'] == 'system' %}{{ bos_token + messages[0]['content'] + '\n' }}"
        "{% set messages = messages[1:] %}{% endif %}"
        "{% for message in messages %}"
        "{% if message['role'] == 'user' %}"
        "{% if add_stream_query_prompt %}{{ ']\nUser: ' + message['content'] }}"
        "{% else %}{{ '\nUser: ' + message['content'] }}{% endif %}"
        "{% elif message['role'] == 'assistant' %}{{ '\nAssistant: '  + message['content'] + eos_token }}"
        "{% elif message['role'] == 'stream' and message['num_frames'] > 0: %}{{ '\n[' + " + stream_placeholder_jinja2 + " + ']' }}"
        "{% endif %}{% endfor %}"
        "{% if add_generation_prompt %}{{ '\nAssistant:' }}"
        "{% elif add_stream_prompt %}{{ '\n[' }}"
        "{% elif add_stream_generation_prompt %}{{ ']\nAssistant:' }}{% endif %}"
    )
