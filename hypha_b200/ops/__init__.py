"""sm_100a op library: differentiable wrappers (functional) + native loader."""
from .functional import (attention_bwd_tcgen05, attention_fwd_tcgen05, set_attention_fwd_version, set_attention_fwd_poly, fallback_counts, reset_fallback_counts, interleave_gate_up, split_gate_up, E4M3, E5M2, bump_fp8_weight_version, fp8_enabled, gemm_fp8, quantize_fp8, set_fp8, fp8_mode, quantize_mxfp8, gemm_mxfp8, mxfp8_quantize_ref, mxfp8_sf_to_table, mxfp8_dequant_ref, mx_sf_shape,
                         add_bias, add_norm_linear, add_rmsnorm, attention_qkv, get_attention_backend,
                         set_attention_backend, swiglu_linear, cross_entropy_inplace, embedding, gemm,
                         get_gemm_backend, linear, linear_cross_entropy, qk_norm, rmsnorm, rope_qkv,
                         rope_tables, set_gemm_backend, swiglu)
from .native import launch_count, load_C, load_core, reset_launch_count

__all__ = [
    "attention_fwd_tcgen05", "attention_bwd_tcgen05", "set_attention_fwd_version", "set_attention_fwd_poly", "fallback_counts", "reset_fallback_counts", "interleave_gate_up", "split_gate_up",
    "E4M3", "E5M2", "bump_fp8_weight_version", "fp8_enabled", "gemm_fp8", "quantize_fp8", "set_fp8", "fp8_mode", "quantize_mxfp8", "gemm_mxfp8", "mxfp8_quantize_ref",
    "mxfp8_sf_to_table", "mxfp8_dequant_ref", "mx_sf_shape",
    "add_bias", "add_norm_linear", "add_rmsnorm", "attention_qkv", "get_attention_backend", "set_attention_backend",
    "swiglu_linear", "cross_entropy_inplace", "embedding", "gemm",
    "get_gemm_backend", "linear", "linear_cross_entropy", "qk_norm", "rmsnorm", "rope_qkv", "rope_tables",
    "set_gemm_backend", "swiglu", "launch_count", "load_C", "load_core", "reset_launch_count",
]
