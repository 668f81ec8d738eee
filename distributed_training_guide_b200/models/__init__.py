"""Model factory (reference: ``AutoConfig.from_pretrained`` + ``AutoModelForCausalLM.from_config``,
``01-single-gpu/train_llm.py:46-49``): random-init causal LMs resolved from the embedded
config registry, in train mode, on a real device or the ``meta`` device."""
from __future__ import annotations

import torch

from .configs import REGISTRY, ModelConfig, get_config, to_hf_config_dict
from .gpt2 import GPT2LMHeadModel
from .llama import LlamaDecoderLayer, LlamaForCausalLM, build_llama

__all__ = ["ModelConfig", "get_config", "build_model", "LlamaForCausalLM", "GPT2LMHeadModel",
           "LlamaDecoderLayer", "REGISTRY", "to_hf_config_dict"]


def build_model(config, dtype=torch.bfloat16, device=None, tp_size=1, init=True):
    """Instantiate (and unless ``init=False`` or on ``meta``, randomly initialise) the model."""
    if isinstance(config, str):
        config = get_config(config)
    if config.arch == "gpt2":
        assert tp_size == 1, "tensor parallelism is implemented for the Llama family"
        model = GPT2LMHeadModel(config, dtype=dtype, device=device)
    else:
        model = LlamaForCausalLM(config, dtype=dtype, device=device, tp_size=tp_size)
    is_meta = device is not None and torch.device(device).type == "meta"
    if init and not is_meta:
        model.init_weights()
    model.train()
    return model
