"""Embedded model configurations.

The GPU boxes have no network, so the handful of Hugging Face model ids the
guide uses (reference: every chapter's ``-m/--model-name`` flag, e.g.
``02-distributed-data-parallel/train_llm.py:57``) are resolved from this table
instead of the hub.  A local directory containing a ``config.json`` is also
accepted, and tiny ``debug-*`` configs exist for tests.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Optional


@dataclasses.dataclass
class ModelConfig:
    arch: str  # "llama" | "gpt2"
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    max_position_embeddings: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    rope_scaling: Optional[dict] = None
    tie_word_embeddings: bool = False
    # gpt2 only
    layer_norm_epsilon: float = 1e-5
    dropout: float = 0.0
    name: str = ""

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def num_parameters(self) -> int:
        h, i, v, l = self.hidden_size, self.intermediate_size, self.vocab_size, self.num_hidden_layers
        if self.arch == "gpt2":
            per_layer = (3 * h * h + 3 * h) + (h * h + h) + (h * i + i) + (i * h + h) + 4 * h
            return v * h + self.max_position_embeddings * h + l * per_layer + 2 * h
        kv = self.num_key_value_heads * self.head_dim
        per_layer = h * h + 2 * kv * h + h * h + 3 * h * i + 2 * h
        n = v * h + l * per_layer + h
        if not self.tie_word_embeddings:
            n += v * h
        return n

    def to_dict(self) -> dict:
        return dataclasses.asdict(self)


_LLAMA3_SCALING = {
    "rope_type": "llama3",
    "factor": 8.0,
    "low_freq_factor": 1.0,
    "high_freq_factor": 4.0,
    "original_max_position_embeddings": 8192,
}


def _llama(name, v, h, i, l, nh, nkv, maxpos, theta, scaling=None):
    return ModelConfig(
        arch="llama", vocab_size=v, hidden_size=h, intermediate_size=i, num_hidden_layers=l,
        num_attention_heads=nh, num_key_value_heads=nkv, max_position_embeddings=maxpos,
        rms_norm_eps=1e-5, rope_theta=theta, rope_scaling=scaling, name=name,
    )


_GPT2 = ModelConfig(
    arch="gpt2", vocab_size=50257, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
    num_attention_heads=12, num_key_value_heads=12, max_position_embeddings=1024,
    tie_word_embeddings=True, layer_norm_epsilon=1e-5, dropout=0.1, name="openai-community/gpt2",
)

REGISTRY = {
    "openai-community/gpt2": _GPT2,
    "gpt2": _GPT2,
    "meta-llama/Llama-2-7b-hf": _llama("meta-llama/Llama-2-7b-hf", 32000, 4096, 11008, 32, 32, 32, 4096, 1e4),
    "meta-llama/Llama-2-13b-hf": _llama("meta-llama/Llama-2-13b-hf", 32000, 5120, 13824, 40, 40, 40, 4096, 1e4),
    "meta-llama/Meta-Llama-3-8B": _llama("meta-llama/Meta-Llama-3-8B", 128256, 4096, 14336, 32, 32, 8, 8192, 5e5),
    "meta-llama/Llama-3.1-8B": _llama("meta-llama/Llama-3.1-8B", 128256, 4096, 14336, 32, 32, 8, 131072, 5e5, _LLAMA3_SCALING),
    "meta-llama/Meta-Llama-3-70B": _llama("meta-llama/Meta-Llama-3-70B", 128256, 8192, 28672, 80, 64, 8, 8192, 5e5),
    "meta-llama/Llama-3.1-70B": _llama("meta-llama/Llama-3.1-70B", 128256, 8192, 28672, 80, 64, 8, 131072, 5e5, _LLAMA3_SCALING),
    "meta-llama/Llama-3.1-405B": _llama("meta-llama/Llama-3.1-405B", 128256, 16384, 53248, 126, 128, 8, 131072, 5e5, _LLAMA3_SCALING),
    "meta-llama/Meta-Llama-3.1-405B": _llama("meta-llama/Meta-Llama-3.1-405B", 128256, 16384, 53248, 126, 128, 8, 131072, 5e5, _LLAMA3_SCALING),
    # tiny configs for tests / smoke runs (head_dim 128 so the sm_100a attention kernel applies)
    "debug-llama": _llama("debug-llama", 1024, 256, 512, 2, 2, 2, 2048, 1e4),
    "debug-llama-gqa": _llama("debug-llama-gqa", 1024, 512, 1024, 2, 4, 2, 2048, 5e5),
    "debug-llama-tp": _llama("debug-llama-tp", 2048, 1024, 2048, 2, 8, 8, 2048, 1e4),
    "debug-gpt2": dataclasses.replace(_GPT2, vocab_size=512, hidden_size=64, intermediate_size=256,
                                      num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                                      max_position_embeddings=128, name="debug-gpt2"),
}


def _from_hf_dict(d: dict, name: str) -> ModelConfig:
    mt = d.get("model_type", "llama")
    if mt == "gpt2":
        h = d.get("n_embd", 768)
        return ModelConfig(
            arch="gpt2", vocab_size=d.get("vocab_size", 50257), hidden_size=h,
            intermediate_size=d.get("n_inner") or 4 * h, num_hidden_layers=d.get("n_layer", 12),
            num_attention_heads=d.get("n_head", 12), num_key_value_heads=d.get("n_head", 12),
            max_position_embeddings=d.get("n_positions", 1024), tie_word_embeddings=True,
            layer_norm_epsilon=d.get("layer_norm_epsilon", 1e-5), dropout=d.get("resid_pdrop", 0.1), name=name,
        )
    if mt != "llama":
        raise ValueError(f"unsupported model_type {mt!r} in {name}")
    scaling = d.get("rope_scaling")
    theta = d.get("rope_theta", 1e4)
    if isinstance(d.get("rope_parameters"), dict):  # transformers>=5 layout
        rp = d["rope_parameters"]
        theta = rp.get("rope_theta", theta)
        if rp.get("rope_type", "default") != "default":
            scaling = rp
    return ModelConfig(
        arch="llama", vocab_size=d["vocab_size"], hidden_size=d["hidden_size"],
        intermediate_size=d["intermediate_size"], num_hidden_layers=d["num_hidden_layers"],
        num_attention_heads=d["num_attention_heads"],
        num_key_value_heads=d.get("num_key_value_heads", d["num_attention_heads"]),
        max_position_embeddings=d.get("max_position_embeddings", 4096),
        rms_norm_eps=d.get("rms_norm_eps", 1e-5), rope_theta=theta, rope_scaling=scaling,
        tie_word_embeddings=d.get("tie_word_embeddings", False), name=name,
    )


def get_config(name: str, **overrides) -> ModelConfig:
    """Resolve ``name`` (registry id, or a directory / file holding an HF ``config.json``)."""
    if name in REGISTRY:
        cfg = REGISTRY[name]
    else:
        path = name
        if os.path.isdir(path):
            path = os.path.join(path, "config.json")
        if not os.path.isfile(path):
            raise KeyError(
                f"unknown model {name!r}: not in the embedded registry ({sorted(REGISTRY)}) "
                "and not a local directory with a config.json (there is no network on the GPU box)"
            )
        with open(path) as fp:
            cfg = _from_hf_dict(json.load(fp), name)
    if overrides:
        cfg = dataclasses.replace(cfg, **overrides)
    return cfg


def to_hf_config_dict(cfg: ModelConfig) -> dict:
    """An HF-style ``config.json`` payload (used to feed the *reference* scripts offline)."""
    if cfg.arch == "gpt2":
        return {
            "model_type": "gpt2", "architectures": ["GPT2LMHeadModel"], "vocab_size": cfg.vocab_size,
            "n_embd": cfg.hidden_size, "n_inner": cfg.intermediate_size, "n_layer": cfg.num_hidden_layers,
            "n_head": cfg.num_attention_heads, "n_positions": cfg.max_position_embeddings,
            "layer_norm_epsilon": cfg.layer_norm_epsilon, "resid_pdrop": cfg.dropout,
            "embd_pdrop": cfg.dropout, "attn_pdrop": cfg.dropout, "activation_function": "gelu_new",
        }
    d = {
        "model_type": "llama", "architectures": ["LlamaForCausalLM"], "vocab_size": cfg.vocab_size,
        "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
        "num_hidden_layers": cfg.num_hidden_layers, "num_attention_heads": cfg.num_attention_heads,
        "num_key_value_heads": cfg.num_key_value_heads, "max_position_embeddings": cfg.max_position_embeddings,
        "rms_norm_eps": cfg.rms_norm_eps, "rope_theta": cfg.rope_theta, "hidden_act": "silu",
        "tie_word_embeddings": cfg.tie_word_embeddings, "attention_bias": False, "mlp_bias": False,
        "bos_token_id": 1, "eos_token_id": 2, "torch_dtype": "bfloat16",
    }
    if cfg.rope_scaling:
        d["rope_scaling"] = cfg.rope_scaling
    return d
