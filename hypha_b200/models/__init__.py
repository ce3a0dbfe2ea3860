"""Model families: Llama (sm_100a kernels), GPT-2 and LeNet (CPU plumbing / reference defaults)."""
from .llama import PRESETS, FlatParams, LlamaConfig, LlamaForCausalLM, get_config

__all__ = ["PRESETS", "FlatParams", "LlamaConfig", "LlamaForCausalLM", "get_config"]
