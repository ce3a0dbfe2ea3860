"""Differentiable front-ends of the sm_100a kernels.

Every op has two implementations:

* CUDA tensors  -> the hand-written kernels in ``hypha_b200/_C`` (bf16 I/O);
* CPU tensors   -> a plain PyTorch implementation of the same math, used by the
  CPU plumbing configuration (GPT-2 / gloo) and as the numerics reference in
  ``tests/``.

Weight gradients follow the Megatron ``main_grad`` convention: when a parameter
carries a ``main_grad`` view (a slice of the model's flat gradient buffer) the
backward kernels write into it directly -- the wgrad GEMM epilogue either
overwrites or accumulates depending on ``param.grad_accumulate`` -- and autograd
sees ``None``.  This is what lets one flat fused-AdamW launch update the model.

Reference parity: these are the ops the reference reaches through
``outputs = model(**batch)`` / ``accelerator.backward(loss)``
(executors/accelerate/src/hypha/accelerate_executor/training.py:108-112).
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .native import cuda_stream, load_C

_GEMM_BACKEND = os.environ.get("HYPHA_GEMM", "tcgen05")


def set_gemm_backend(name: str) -> None:
    """``tcgen05`` (hand-written kernel, default) or ``cublas`` (torch.mm, baseline)."""
    global _GEMM_BACKEND
    if name not in ("tcgen05", "cublas"):
        raise ValueError(name)
    _GEMM_BACKEND = name


def get_gemm_backend() -> str:
    return _GEMM_BACKEND


# Library fallbacks are counted, never silent: a CUDA tensor that reaches torch.mm / cuDNN although the
# hand-written backend was requested bumps a counter (and logs once per reason); bench.py refuses to report a
# number if any happened inside its timed region, and HYPHA_STRICT_KERNELS=1 turns them into errors.
_FALLBACKS: dict = {}
_FALLBACK_LOGGED: set = set()


def _note_fallback(op: str, why: str) -> None:
    key = f"{op}: {why}"
    _FALLBACKS[key] = _FALLBACKS.get(key, 0) + 1
    if os.environ.get("HYPHA_STRICT_KERNELS") == "1":
        raise RuntimeError(f"library fallback refused (HYPHA_STRICT_KERNELS=1): {key}")
    if key not in _FALLBACK_LOGGED:
        _FALLBACK_LOGGED.add(key)
        import logging

        logging.getLogger("hypha.ops").warning("library fallback: %s", key)


def fallback_counts() -> dict:
    """``{"op: reason": count}`` of CUDA calls that ran a library kernel instead of the requested sm_100a one."""
    return dict(_FALLBACKS)


def reset_fallback_counts() -> None:
    _FALLBACKS.clear()


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


# -----------------------------------------------------------------------------
# GEMM
# -----------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_kmajor: bool = True, b_kmajor: bool = True,
         out: Optional[torch.Tensor] = None, accumulate: bool = False,
         swiglu_out: Optional[torch.Tensor] = None, dswiglu: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
         ) -> torch.Tensor:
    """``D[M,N] (+)= A . B`` without materialising transposes.

    ``a_kmajor``: ``a`` is stored ``[M, K]`` else ``[K, M]``;
    ``b_kmajor``: ``b`` is stored ``[N, K]`` else ``[K, N]`` (row-major, unit inner stride).
    ``swiglu_out`` ``[M, N/2]``: the GEMM epilogue also writes ``silu(D[:, 2i]) * D[:, 2i+1]`` there (fused
    SwiGLU on interleaved gate/up columns; tcgen05 2-CTA kernel, N % 128 == 0).
    ``dswiglu = (gu, dgu)``: SwiGLU BACKWARD epilogue -- the product is d(act) (rounded to bf16, not stored unless
    the fallback runs) and ``dgu[M, 2N] = swiglu'(gu) * d(act)`` is written instead; returns ``dgu``.
    """
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
    N, Kb = (b.shape[0], b.shape[1]) if b_kmajor else (b.shape[1], b.shape[0])
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
        accumulate = False
    assert out.shape == (M, N) and out.stride(1) == 1
    use_native = (a.is_cuda and _GEMM_BACKEND == "tcgen05" and a.dtype == torch.bfloat16
                  and b.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
                  and N % 8 == 0 and K % 8 == 0 and (a_kmajor or M % 8 == 0)
                  and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0 and out.stride(0) % 8 == 0
                  and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
                  and out.data_ptr() % 16 == 0)
    if dswiglu is not None:
        gu_t, dgu_t = dswiglu
        if (use_native and N % 64 == 0 and not accumulate and gu_t.stride(1) == 1 and gu_t.stride(0) % 8 == 0
                and dgu_t.stride(1) == 1 and dgu_t.stride(0) % 8 == 0 and gu_t.data_ptr() % 16 == 0
                and dgu_t.data_ptr() % 16 == 0):
            load_C().gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0),
                               out.stride(0), int(a_kmajor), int(b_kmajor), 0, cuda_stream(), dgu_t.data_ptr(),
                               dgu_t.stride(0), gu_t.data_ptr(), gu_t.stride(0))
            return dgu_t
        dact = gemm(a, b, a_kmajor=a_kmajor, b_kmajor=b_kmajor, out=out, accumulate=accumulate)
        if dact.is_cuda:
            load_C().swiglu_bwd(dact.data_ptr(), gu_t.data_ptr(), dgu_t.data_ptr(), M, N, cuda_stream())
        else:
            raise RuntimeError("dswiglu epilogue is a CUDA path")
        return dgu_t
    fuse = swiglu_out is not None and use_native and N % 128 == 0 and not accumulate
    if use_native:
        load_C().gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0),
                           b.stride(0), out.stride(0), int(a_kmajor), int(b_kmajor),
                           int(accumulate), cuda_stream(), swiglu_out.data_ptr() if fuse else 0,
                           swiglu_out.stride(0) if fuse else 0)
        if swiglu_out is not None and not fuse:
            load_C().swiglu_fwd(out.data_ptr(), swiglu_out.data_ptr(), M, N // 2, cuda_stream())
        return out
    if a.is_cuda and _GEMM_BACKEND == "tcgen05":
        _note_fallback("gemm", f"operands not 16-byte aligned / multiple of 8 (M={M} N={N} K={K}, dtypes "
                               f"{a.dtype}/{b.dtype}/{out.dtype})")
    A = a if a_kmajor else a.t()
    B = b.t() if b_kmajor else b
    if accumulate:
        out.addmm_(A, B)
    else:
        torch.mm(A, B, out=out)
    if swiglu_out is not None:
        swiglu_out.copy_(swiglu(out))
    return out


# -----------------------------------------------------------------------------
# fp8 paths.  "mx" (default): MXFP8 -- e4m3 values, one UE8M0 scale per 32 elements along the contraction
# dimension, applied inside the tensor core (tcgen05.mma.kind::mxf8f6f4.block_scale).  "tensor": per-tensor current
# scaling on kind::f8f6f4 (e4m3 forward operands, e5m2 gradients), the round-1 path.
# -----------------------------------------------------------------------------
_FP8 = os.environ.get("HYPHA_FP8", "0") == "1"
_FP8_MODE = os.environ.get("HYPHA_FP8_MODE", "mx")
_FP8_WEIGHT_VERSION = 0
E4M3, E5M2 = 0, 1


def set_fp8(enabled: bool, mode: Optional[str] = None) -> None:
    """Run Linear-layer GEMMs (fwd, dgrad, wgrad) in fp8 on the tcgen05 kernels: ``mode="mx"`` block-scaled MXFP8
    (``kind::mxf8f6f4.block_scale``), ``mode="tensor"`` per-tensor scaling (``kind::f8f6f4``).
    Master weights, gradients written to ``main_grad``, the optimizer and the DiLoCo outer step stay
    bf16/fp32 (BASELINE config 3: fp8 inner steps, bf16 outer all-reduce)."""
    global _FP8, _FP8_MODE
    _FP8 = bool(enabled)
    if mode is not None:
        if mode not in ("mx", "tensor"):
            raise ValueError(mode)
        _FP8_MODE = mode


def fp8_mode() -> str:
    return _FP8_MODE


def fp8_enabled() -> bool:
    return _FP8


def bump_fp8_weight_version() -> None:
    """Called by the optimizer after it rewrote the bf16 weights: cached fp8 copies are stale."""
    global _FP8_WEIGHT_VERSION
    _FP8_WEIGHT_VERSION += 1


def quantize_fp8(x: torch.Tensor, fmt: int = E4M3) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(q, dequant)``: ``q = fp8(x * fmax/amax(x))`` as uint8, ``dequant = amax/fmax`` (device fp32[1])."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0
    C = load_C()
    s = cuda_stream()
    st = torch.empty(2, dtype=torch.float32, device=x.device)
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    C.amax_bf16(x.data_ptr(), x.numel(), st.data_ptr(), s)
    C.quantize_fp8(x.data_ptr(), q.data_ptr(), x.numel(), st.data_ptr(), st.data_ptr() + 4, fmt, s)
    return q, st[1:2]


_FP8_CACHE_WEIGHTS = os.environ.get("HYPHA_FP8_CACHE", "1") == "1"


def _weight_fp8(weight: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp8 copy of a weight; cached until the optimizer rewrites the weights (costs 1 byte/param;
    HYPHA_FP8_CACHE=0 re-quantises on every use instead)."""
    if not _FP8_CACHE_WEIGHTS:
        return quantize_fp8(weight.detach(), E4M3)
    cache = getattr(weight, "_fp8_cache", None)
    if cache is not None and cache[0] == _FP8_WEIGHT_VERSION:
        return cache[1], cache[2]
    q, dq = quantize_fp8(weight.detach(), E4M3)
    weight._fp8_cache = (_FP8_WEIGHT_VERSION, q, dq)
    return q, dq


def gemm_fp8(aq: torch.Tensor, a_dq: torch.Tensor, bq: torch.Tensor, b_dq: torch.Tensor, *, a_kmajor: bool = True,
             b_kmajor: bool = True, a_fmt: int = E4M3, b_fmt: int = E4M3, out: Optional[torch.Tensor] = None,
             accumulate: bool = False, swiglu_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``D (+)= dq_a*dq_b * (A . B)`` with uint8-stored fp8 operands -> bf16 (``swiglu_out``: see `gemm`)."""
    M, K = (aq.shape[0], aq.shape[1]) if a_kmajor else (aq.shape[1], aq.shape[0])
    N, Kb = (bq.shape[0], bq.shape[1]) if b_kmajor else (bq.shape[1], bq.shape[0])
    assert K == Kb
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=aq.device)
        accumulate = False
    load_C().gemm_fp8(aq.data_ptr(), bq.data_ptr(), out.data_ptr(), M, N, K, aq.stride(0), bq.stride(0),
                      out.stride(0), int(a_kmajor), int(b_kmajor), a_fmt, b_fmt, a_dq.data_ptr(), b_dq.data_ptr(),
                      int(accumulate), cuda_stream(),
                      swiglu_out.data_ptr() if (swiglu_out is not None and N % 128 == 0 and not accumulate) else 0,
                      swiglu_out.stride(0) if swiglu_out is not None else 0)
    if swiglu_out is not None and not (N % 128 == 0 and not accumulate):
        load_C().swiglu_fwd(out.data_ptr(), swiglu_out.data_ptr(), M, N // 2, cuda_stream())
    return out


# ---- MXFP8 ---------------------------------------------------------------------------------------------------
def mx_sf_shape(rows: int, k: int) -> Tuple[int, int, int]:
    """Shape of the scale table of a ``[rows, k]`` operand: one 512-byte atom per (128 K-elements, 128 rows),
    K blocks outermost."""
    return (k // 128, rows // 128, 512)


def quantize_mxfp8(x: torch.Tensor, rowwise: bool = True, transposed: bool = False):
    """``(q, sf, qT, sfT)`` of a bf16 matrix ``x[R, C]`` in ONE pass (`csrc/kernels/fp8.cu`):

    ``q[R, C]`` e4m3 (uint8) scaled per 32 elements along ``C`` with table ``sf`` -- the operand of a GEMM that
    contracts over ``C``; ``qT[C, R]`` = ``x^T`` scaled along ``R`` with ``sfT`` -- the operand when the GEMM
    contracts over ``R`` (wgrad, dgrad's weight).  Pairs not requested are ``None``.  ``R, C % 128 == 0``."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    q = sf = qT = sfT = None
    if rowwise:
        q = torch.empty((R, C), dtype=torch.uint8, device=x.device)
        sf = torch.empty(mx_sf_shape(R, C), dtype=torch.uint8, device=x.device)
    if transposed:
        qT = torch.empty((C, R), dtype=torch.uint8, device=x.device)
        sfT = torch.empty(mx_sf_shape(C, R), dtype=torch.uint8, device=x.device)
    load_C().quantize_mxfp8(x.data_ptr(), x.stride(0), _ptr(q), C, _ptr(sf), _ptr(qT), R, _ptr(sfT), R, C, cuda_stream())
    return q, sf, qT, sfT


def mxfp8_quantize_ref(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Plain-torch model of the row-wise quantiser: ``(q as float8_e4m3fn [R, C], biased exponents uint8 [R, C/32])``."""
    R, C = x.shape
    xb = x.float().reshape(R, C // 32, 32)
    amax = xb.abs().amax(dim=-1)
    u = (amax * (1.0 / 448.0)).view(torch.int32)
    e = ((u >> 23) & 0xFF) + ((u & 0x7FFFFF) != 0).to(torch.int32)
    e = e.clamp(max=254)
    inv = ((254 - e) << 23).view(torch.float32)
    q = (xb * inv[..., None]).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).reshape(R, C)
    return q, e.to(torch.uint8)


def mxfp8_sf_to_table(e: torch.Tensor) -> torch.Tensor:
    """Biased exponents ``[rows, k/32]`` -> the atom-layout table ``[k/128, rows/128, 512]`` the kernels use."""
    rows, kb = e.shape
    t = e.reshape(rows // 128, 4, 32, kb // 4, 4)        # [row block, r/32, r%32, k block, k/32]
    return t.permute(3, 0, 2, 1, 4).reshape(kb // 4, rows // 128, 512).contiguous()


def mxfp8_dequant_ref(q: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """fp32 matrix an MXFP8 operand stands for (uint8-stored e4m3 + atom-layout scale table)."""
    rows, k = q.shape
    e = table.reshape(k // 128, rows // 128, 32, 4, 4).permute(1, 3, 2, 0, 4).reshape(rows, k // 32)
    scale = torch.exp2(e.float() - 127.0)
    return (q.view(torch.float8_e4m3fn).float().reshape(rows, k // 32, 32) * scale[..., None]).reshape(rows, k)


def gemm_mxfp8(aq: torch.Tensor, asf: torch.Tensor, bq: torch.Tensor, bsf: torch.Tensor, *,
               out: Optional[torch.Tensor] = None, accumulate: bool = False, swiglu_out: Optional[torch.Tensor] = None,
               dswiglu: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """``D[M,N] (+)= A[M,K] . B[N,K]^T`` with MXFP8 operands (both K-major, scale tables from `quantize_mxfp8`);
    block scales applied by ``tcgen05.mma.kind::mxf8f6f4.block_scale``, fp32 accumulation, bf16 output.
    ``swiglu_out`` / ``dswiglu``: the fused epilogues of `gemm`."""
    M, K = aq.shape
    N, Kb = bq.shape
    assert K == Kb and aq.stride(1) == 1 and bq.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=aq.device)
        accumulate = False
    act, ld_act, gu, ld_gu = 0, 0, 0, 0
    if dswiglu is not None:
        gu_t, dgu_t = dswiglu
        act, ld_act, gu, ld_gu = dgu_t.data_ptr(), dgu_t.stride(0), gu_t.data_ptr(), gu_t.stride(0)
    elif swiglu_out is not None:
        act, ld_act = swiglu_out.data_ptr(), swiglu_out.stride(0)
    load_C().gemm_mxfp8(aq.data_ptr(), asf.data_ptr(), bq.data_ptr(), bsf.data_ptr(), out.data_ptr(), M, N, K,
                        aq.stride(0), bq.stride(0), out.stride(0), E4M3, E4M3, int(accumulate), cuda_stream(),
                        act, ld_act, gu, ld_gu)
    return dswiglu[1] if dswiglu is not None else out


_MX_CACHE_WEIGHTS = os.environ.get("HYPHA_MX_CACHE", "0") == "1"


def _weight_mx(weight: torch.Tensor, transposed: bool):
    """MXFP8 operand of a weight ``[N, K]``: ``(q, sf)`` for the forward GEMM, or ``(qT, sfT)`` (W^T scaled along N)
    for dgrad.  Quantised where it is used: one read of the bf16 weight and one fp8 write per use is ~1 % of the
    step, while keeping both copies resident costs 2 bytes / parameter -- 16 GB on the 8B model, i.e. more layers to
    recompute.  ``HYPHA_MX_CACHE=1`` keeps them until the optimizer rewrites the weights."""
    if _MX_CACHE_WEIGHTS:
        cache = getattr(weight, "_mx_cache", None)
        if cache is None or cache[0] != _FP8_WEIGHT_VERSION:
            cache = weight._mx_cache = (_FP8_WEIGHT_VERSION, quantize_mxfp8(weight.detach(), rowwise=True, transposed=True))
        quad = cache[1]
        return (quad[2], quad[3]) if transposed else (quad[0], quad[1])
    quad = quantize_mxfp8(weight.detach(), rowwise=not transposed, transposed=transposed)
    return (quad[2], quad[3]) if transposed else (quad[0], quad[1])


def _mx_ok(x2: torch.Tensor, weight: torch.Tensor) -> bool:
    return (_FP8 and _FP8_MODE == "mx" and x2.is_cuda and x2.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
            and x2.shape[0] % 128 == 0 and x2.shape[1] % 128 == 0 and weight.shape[0] % 128 == 0
            and x2.stride(1) == 1 and x2.stride(0) % 8 == 0 and weight.is_contiguous())


def _fp8_ok(x2: torch.Tensor, weight: torch.Tensor) -> bool:
    if _FP8 and _FP8_MODE == "mx":
        return _mx_ok(x2, weight)
    return (_FP8 and x2.is_cuda and x2.dtype == torch.bfloat16 and x2.shape[0] % 16 == 0
            and x2.shape[1] % 16 == 0 and weight.shape[0] % 16 == 0)


def _mm_fwd(x2: torch.Tensor, weight: torch.Tensor, swiglu_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x2 @ weight.T on the bf16 or fp8 tensor-core path (``swiglu_out``: fused SwiGLU epilogue, see `gemm`)."""
    if _mx_ok(x2, weight):
        xq, xsf, _, _ = quantize_mxfp8(x2, rowwise=True, transposed=False)
        wq, wsf = _weight_mx(weight, transposed=False)
        return gemm_mxfp8(xq, xsf, wq, wsf, swiglu_out=swiglu_out)
    if _fp8_ok(x2, weight):
        xq, xdq = quantize_fp8(x2, E4M3)
        wq, wdq = _weight_fp8(weight)
        return gemm_fp8(xq, xdq, wq, wdq, swiglu_out=swiglu_out)
    return gemm(x2, weight, swiglu_out=swiglu_out)


def _mm_bwd(dy2: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor, need_dx: bool = True,
            dx_out: Optional[torch.Tensor] = None, dswiglu: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """(dx, dW) for y = x2 @ weight.T.  dW goes to ``weight.main_grad`` when present (returns None).
    ``dswiglu = (gu, dgu)`` (MXFP8 only): the dgrad epilogue writes d(gate|up) into ``dgu`` and returns it as dx."""
    mg = getattr(weight, "main_grad", None)
    acc = bool(getattr(weight, "grad_accumulate", False)) if mg is not None else False
    if _mx_ok(x2, weight):
        # every operand K-major: dW[N,K] = dy^T[N,T] . (x^T[K,T])^T, dx[T,K] = dy[T,N] . (W^T[K,N])^T
        dyq, dysf, dyTq, dyTsf = quantize_mxfp8(dy2, rowwise=need_dx, transposed=True)
        _, _, xTq, xTsf = quantize_mxfp8(x2, rowwise=False, transposed=True)
        dw = gemm_mxfp8(dyTq, dyTsf, xTq, xTsf, out=mg, accumulate=acc)
        dx = None
        if need_dx:
            wTq, wTsf = _weight_mx(weight, transposed=True)
            dx = gemm_mxfp8(dyq, dysf, wTq, wTsf, out=dx_out, dswiglu=dswiglu)
        return dx, (None if mg is not None else dw)
    assert dswiglu is None, "the fused SwiGLU-backward epilogue of _mm_bwd is an MXFP8 path"
    if _fp8_ok(x2, weight) and dy2.shape[1] % 16 == 0:
        dyq, dydq = quantize_fp8(dy2, E5M2)
        xq, xdq = quantize_fp8(x2, E4M3)
        dw = gemm_fp8(dyq, dydq, xq, xdq, a_kmajor=False, b_kmajor=False, a_fmt=E5M2, b_fmt=E4M3, out=mg,
                      accumulate=acc)
        dx = None
        if need_dx:
            wq, wdq = _weight_fp8(weight)
            dx = gemm_fp8(dyq, dydq, wq, wdq, a_kmajor=True, b_kmajor=False, a_fmt=E5M2, b_fmt=E4M3, out=dx_out)
        return dx, (None if mg is not None else dw)
    dw = _wgrad_into(weight, dy2, x2)
    dx = gemm(dy2, weight, a_kmajor=True, b_kmajor=False, out=dx_out) if need_dx else None
    return dx, dw


def _wgrad_into(param: torch.Tensor, dy2: torch.Tensor, x2: torch.Tensor) -> Optional[torch.Tensor]:
    """dW[N,K] = dy^T . x ; written into ``param.main_grad`` when present."""
    mg = getattr(param, "main_grad", None)
    if mg is not None:
        gemm(dy2, x2, a_kmajor=False, b_kmajor=False, out=mg,
             accumulate=bool(getattr(param, "grad_accumulate", False)))
        return None
    return gemm(dy2, x2, a_kmajor=False, b_kmajor=False)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        x2 = x.reshape(-1, x.shape[-1])
        y = _mm_fwd(x2, weight)
        ctx.save_for_backward(x2, weight)
        ctx.x_shape = x.shape
        return y if x.dim() == 2 else y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if dy2.stride(1) != 1 or dy2.stride(0) % 8 != 0:
            dy2 = dy2.contiguous()
        dx, dw = _mm_bwd(dy2, x2, weight, need_dx=ctx.needs_input_grad[0])
        if dx is not None:
            dx = dx.view(ctx.x_shape)
        return dx, dw


def linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``y = x @ weight.T`` (no bias); fwd/dgrad/wgrad all on the tcgen05 GEMM."""
    if not x.is_cuda:
        return F.linear(x, weight)
    return _Linear.apply(x, weight)


# -----------------------------------------------------------------------------
# RMSNorm (+ fused residual add)
# -----------------------------------------------------------------------------
def _rmsnorm_ref(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * weight.float()).to(x.dtype)


def _norm_wgrad(param: torch.Tensor, partial: torch.Tensor, nparts: int, cols: int):
    C = load_C()
    mg = getattr(param, "main_grad", None)
    if mg is not None:
        C.colsum_partials(partial.data_ptr(), nparts, cols, mg.data_ptr(),
                          int(bool(getattr(param, "grad_accumulate", False))), cuda_stream())
        return None
    dw = torch.empty(cols, dtype=torch.bfloat16, device=partial.device)
    C.colsum_partials(partial.data_ptr(), nparts, cols, dw.data_ptr(), 0, cuda_stream())
    return dw.to(param.dtype)


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        C = load_C()
        cols = x.shape[-1]
        x2 = x.reshape(-1, cols)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        C.rmsnorm_fwd(x2.data_ptr(), 0, 0, weight.data_ptr(), y.data_ptr(), rstd.data_ptr(), rows,
                      cols, eps, cuda_stream())
        ctx.save_for_backward(x2, weight, rstd)
        ctx.shape = x.shape
        return y if x.dim() == 2 else y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        C = load_C()
        x2, weight, rstd = ctx.saved_tensors
        rows, cols = x2.shape
        dy2 = dy.reshape(rows, cols).contiguous()
        dx = torch.empty_like(x2)
        nparts = C.rmsnorm_bwd_num_partials(rows)
        partial = torch.empty((nparts, cols), dtype=torch.float32, device=x2.device)
        C.rmsnorm_bwd(dy2.data_ptr(), x2.data_ptr(), weight.data_ptr(), rstd.data_ptr(), 0,
                      dx.data_ptr(), partial.data_ptr(), rows, cols, cuda_stream())
        dw = _norm_wgrad(weight, partial, nparts, cols)
        return dx.view(ctx.shape), dw, None


class _AddRMSNorm(torch.autograd.Function):
    """h = x + residual ; y = rmsnorm(h) * w ; returns (y, h)."""

    @staticmethod
    def forward(ctx, x, residual, weight, eps):
        C = load_C()
        cols = x.shape[-1]
        x2 = x.reshape(-1, cols).contiguous()
        r2 = residual.reshape(-1, cols).contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        h = torch.empty_like(x2)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        C.rmsnorm_fwd(x2.data_ptr(), r2.data_ptr(), h.data_ptr(), weight.data_ptr(), y.data_ptr(),
                      rstd.data_ptr(), rows, cols, eps, cuda_stream())
        ctx.save_for_backward(h, weight, rstd)
        ctx.shape = x.shape
        if x.dim() == 2:
            return y, h
        return y.view(x.shape), h.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dh):
        C = load_C()
        h, weight, rstd = ctx.saved_tensors
        rows, cols = h.shape
        dy2 = dy.reshape(rows, cols).contiguous()
        dh2 = dh.reshape(rows, cols).contiguous() if dh is not None else None
        dx = torch.empty_like(h)
        nparts = C.rmsnorm_bwd_num_partials(rows)
        partial = torch.empty((nparts, cols), dtype=torch.float32, device=h.device)
        C.rmsnorm_bwd(dy2.data_ptr(), h.data_ptr(), weight.data_ptr(), rstd.data_ptr(), _ptr(dh2),
                      dx.data_ptr(), partial.data_ptr(), rows, cols, cuda_stream())
        dw = _norm_wgrad(weight, partial, nparts, cols)
        dxv = dx.view(ctx.shape)
        return dxv, dxv, dw, None


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    if not x.is_cuda:
        return _rmsnorm_ref(x, weight, eps)
    return _RMSNorm.apply(x, weight, eps)


def add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused ``h = x + residual; y = rmsnorm(h)``; returns ``(y, h)``."""
    if not x.is_cuda:
        h = (x.float() + residual.float()).to(x.dtype)
        return _rmsnorm_ref(h, weight, eps), h
    return _AddRMSNorm.apply(x, residual, weight, eps)


def _norm_linear_fwd(x, residual, norm_w, weight, eps, act_out=None):
    C = load_C()
    rows, cols = x.shape
    n = torch.empty_like(x)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    if residual is not None:
        h = torch.empty_like(x)
        C.rmsnorm_fwd(x.data_ptr(), residual.data_ptr(), h.data_ptr(), norm_w.data_ptr(), n.data_ptr(),
                      rstd.data_ptr(), rows, cols, eps, cuda_stream())
    else:
        h = x
        C.rmsnorm_fwd(x.data_ptr(), 0, 0, norm_w.data_ptr(), n.data_ptr(), rstd.data_ptr(), rows, cols, eps,
                      cuda_stream())
    return _mm_fwd(n, weight, act_out), h, rstd


def _norm_linear_bwd(dy, dh, h, rstd, norm_w, weight, eps):
    C = load_C()
    rows, cols = h.shape
    s = cuda_stream()
    n = torch.empty_like(h)
    C.rmsnorm_fwd(h.data_ptr(), 0, 0, norm_w.data_ptr(), n.data_ptr(), 0, rows, cols, eps, s)
    dy2 = dy if dy.is_contiguous() else dy.contiguous()
    if _fp8_ok(n, weight):
        dn, dw = _mm_bwd(dy2, n, weight)
    else:
        dn, dw = _mm_bwd(dy2, n, weight, dx_out=n)                           # reuse the buffer
    dx = torch.empty_like(h)
    nparts = C.rmsnorm_bwd_num_partials(rows)
    partial = torch.empty((nparts, cols), dtype=torch.float32, device=h.device)
    dh2 = None
    if dh is not None:
        dh2 = dh if dh.is_contiguous() else dh.contiguous()
    C.rmsnorm_bwd(dn.data_ptr(), h.data_ptr(), norm_w.data_ptr(), rstd.data_ptr(), _ptr(dh2), dx.data_ptr(),
                  partial.data_ptr(), rows, cols, s)
    return dx, _norm_wgrad(norm_w, partial, nparts, cols), dw


class _AddNormLinear(torch.autograd.Function):
    """``h = x + residual; y = rmsnorm(h) @ W^T`` saving only ``h`` and ``rstd``.

    The normalised activation (the GEMM's A operand) is NOT kept: backward re-derives it with one
    extra rmsnorm_fwd launch (memory-bound, ~60 us) instead of holding ``tokens x hidden`` bf16 per
    norm per layer.  Returns ``(y, h)``."""

    @staticmethod
    def forward(ctx, x, residual, norm_w, weight, eps, act_out=None):
        y, h, rstd = _norm_linear_fwd(x, residual, norm_w, weight, eps, act_out)
        ctx.save_for_backward(h, rstd, norm_w, weight)
        ctx.eps = eps
        return y, h

    @staticmethod
    def backward(ctx, dy, dh):
        h, rstd, norm_w, weight = ctx.saved_tensors
        dx, dnw, dw = _norm_linear_bwd(dy, dh, h, rstd, norm_w, weight, ctx.eps)
        return dx, dx, dnw, dw, None, None


class _NormLinear(torch.autograd.Function):
    """``y = rmsnorm(x) @ W^T`` (no residual input); same recompute policy as ``_AddNormLinear``."""

    @staticmethod
    def forward(ctx, x, norm_w, weight, eps, act_out=None):
        y, _, rstd = _norm_linear_fwd(x, None, norm_w, weight, eps, act_out)
        ctx.save_for_backward(x, rstd, norm_w, weight)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, rstd, norm_w, weight = ctx.saved_tensors
        dx, dnw, dw = _norm_linear_bwd(dy, None, x, rstd, norm_w, weight, ctx.eps)
        return dx, dnw, dw, None, None


def add_norm_linear(x: torch.Tensor, residual: Optional[torch.Tensor], norm_w: torch.Tensor,
                    weight: torch.Tensor, eps: float = 1e-5, swiglu_out: Optional[torch.Tensor] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``h = x + residual`` (or ``x``), ``y = rmsnorm(h) * norm_w @ weight.T``; returns ``(y, h)``.

    ``swiglu_out`` (``[rows, out_features/2]``, no autograd): when ``weight`` is an interleaved gate_up
    projection the GEMM epilogue also writes ``silu(gate) * up`` there, which `swiglu_linear(..., act=...)`
    then consumes instead of launching the SwiGLU kernel (the gradient still flows through ``y``)."""
    if not x.is_cuda:
        h = x if residual is None else (x.float() + residual.float()).to(x.dtype)
        y = F.linear(_rmsnorm_ref(h, norm_w, eps), weight)
        if swiglu_out is not None:
            swiglu_out.copy_(swiglu(y.detach()))
        return y, h
    assert x.dim() == 2 and x.is_contiguous()
    if residual is None:
        return _NormLinear.apply(x, norm_w, weight, eps, swiglu_out), x
    return _AddNormLinear.apply(x, residual, norm_w, weight, eps, swiglu_out)


class _SwiGLULinear(torch.autograd.Function):
    """``y = (silu(g) * u) @ W^T`` on a fused gate|up activation, saving only ``gu``."""

    @staticmethod
    def forward(ctx, gu, weight, act=None):
        rows, two_i = gu.shape
        if act is None:            # not produced by the gate_up GEMM's epilogue: one memory-bound launch
            act = torch.empty((rows, two_i // 2), dtype=gu.dtype, device=gu.device)
            load_C().swiglu_fwd(gu.data_ptr(), act.data_ptr(), rows, two_i // 2, cuda_stream())
        y = _mm_fwd(act, weight)
        ctx.save_for_backward(gu, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        C = load_C()
        gu, weight = ctx.saved_tensors
        rows, two_i = gu.shape
        inter = two_i // 2
        s = cuda_stream()
        act = torch.empty((rows, inter), dtype=gu.dtype, device=gu.device)
        C.swiglu_fwd(gu.data_ptr(), act.data_ptr(), rows, inter, s)
        dy2 = dy if dy.is_contiguous() else dy.contiguous()
        dgu = torch.empty_like(gu)
        if _mx_ok(act, weight):
            # one quantisation of dy feeds wgrad and dgrad; the dgrad epilogue turns d(act) into d(gate|up)
            _, dw = _mm_bwd(dy2, act, weight, dx_out=act, dswiglu=(gu, dgu))
        elif _fp8_ok(act, weight):
            dact, dw = _mm_bwd(dy2, act, weight)
            C.swiglu_bwd(dact.data_ptr(), gu.data_ptr(), dgu.data_ptr(), rows, inter, s)
        else:
            # wgrad consumes the recomputed activation; the dgrad GEMM's epilogue then turns its own product
            # d(act) into d(gate|up) on the way out (no [rows, inter] round trip, no SwiGLU-backward launch)
            _, dw = _mm_bwd(dy2, act, weight, need_dx=False)
            gemm(dy2, weight, a_kmajor=True, b_kmajor=False, out=act, dswiglu=(gu, dgu))
        return dgu, dw, None


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """``[2I, H]`` fused weight with rows ``gate_0, up_0, gate_1, up_1, ...`` from HF-style gate / up ``[I, H]``."""
    return torch.stack([gate, up], dim=1).reshape(2 * gate.shape[0], *gate.shape[1:])


def split_gate_up(wgu: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return wgu[0::2], wgu[1::2]


def swiglu_linear(gu: torch.Tensor, weight: torch.Tensor, act: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``(silu(gate) * up) @ weight.T`` with the activation recomputed in backward.  ``act``: the activation
    already computed by the producing GEMM's epilogue (`add_norm_linear(..., swiglu_out=act)`)."""
    if not gu.is_cuda:
        return F.linear(swiglu(gu), weight)
    assert gu.dim() == 2 and gu.is_contiguous()
    return _SwiGLULinear.apply(gu, weight, act)


# -----------------------------------------------------------------------------
# RoPE
# -----------------------------------------------------------------------------
def rope_tables(seq_len: int, head_dim: int, theta: float, device,
                scaling: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(cos, sin)`` tables ``[seq_len, head_dim / 2]`` the RoPE kernel reads.  ``scaling``: the ``llama3`` frequency
    scaling of Llama 3.1+ (long wavelengths divided by ``factor``, short ones untouched, a smooth blend between) --
    it only changes the table, the kernel is the same."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
    if scaling:
        kind = scaling.get("rope_type", scaling.get("type"))
        if kind != "llama3":
            raise ValueError(f"unsupported RoPE scaling {kind!r}")
        factor, lo, hi = float(scaling["factor"]), float(scaling.get("low_freq_factor", 1.0)), float(scaling.get("high_freq_factor", 4.0))
        old_len = float(scaling.get("original_max_position_embeddings", 8192))
        wavelen = 2 * math.pi / inv
        scaled = torch.where(wavelen > old_len / lo, inv / factor, inv)
        smooth = (old_len / wavelen - lo) / (hi - lo)
        blended = (1 - smooth) * scaled / factor + smooth * scaled
        medium = ~(wavelen < old_len / hi) & ~(wavelen > old_len / lo)
        inv = torch.where(medium, blended, scaled)
    ang = torch.outer(torch.arange(seq_len, dtype=torch.float32, device=device), inv)
    return ang.cos().contiguous(), ang.sin().contiguous()


def _rope_ref(qkv, cos, sin, seq_len, n_rot_heads, head_dim):
    T = qkv.shape[0]
    x = qkv.view(T, -1, head_dim)
    rot, rest = x[:, :n_rot_heads].float(), x[:, n_rot_heads:]
    pos = torch.arange(T, device=qkv.device) % seq_len
    c, s = cos[pos][:, None, :], sin[pos][:, None, :]
    x1, x2 = rot[..., : head_dim // 2], rot[..., head_dim // 2:]
    out = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(qkv.dtype)
    return torch.cat([out, rest], dim=1).reshape(T, -1)


class _RoPE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos, sin, seq_len, n_rot_heads, head_dim, grad_inplace):
        assert qkv.dim() == 2 and qkv.stride(1) == 1
        load_C().rope_inplace(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), qkv.shape[0], seq_len,
                              n_rot_heads, head_dim, qkv.stride(0), 0, cuda_stream())
        ctx.mark_dirty(qkv)
        ctx.save_for_backward(cos, sin)
        ctx.args = (seq_len, n_rot_heads, head_dim)
        ctx.grad_inplace = grad_inplace
        return qkv

    @staticmethod
    def backward(ctx, dqkv):
        cos, sin = ctx.saved_tensors
        seq_len, n_rot_heads, head_dim = ctx.args
        if not dqkv.is_contiguous():
            dqkv = dqkv.contiguous()
        elif not ctx.grad_inplace:
            dqkv = dqkv.clone()   # autograd contract: never mutate an incoming gradient
        load_C().rope_inplace(dqkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), dqkv.shape[0],
                              seq_len, n_rot_heads, head_dim, dqkv.stride(0), 1, cuda_stream())
        return dqkv, None, None, None, None, None, None


class _AddBias(torch.autograd.Function):
    """``y = x + bias`` over the rows of ``x [T, N]`` (Qwen2's bias on the fused q|k|v projection).  The bias gradient
    is the fp32 column sum of ``dy``, written to ``bias.main_grad`` under the same overwrite / accumulate convention as
    the weight gradients.  Plain elementwise + reduction kernels: 2 x T x N bytes next to the GEMM that produced x."""

    @staticmethod
    def forward(ctx, x, bias):
        ctx.bias = bias
        return x + bias.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        bias = ctx.bias
        db = dy.reshape(-1, dy.shape[-1]).sum(0, dtype=torch.float32)
        mg = getattr(bias, "main_grad", None)
        if mg is None:
            return dy, db.to(bias.dtype)
        if getattr(bias, "grad_accumulate", False):
            mg.add_(db.to(mg.dtype))
        else:
            mg.copy_(db)
        return dy, None


def add_bias(x: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        return x + bias.to(x.dtype)          # CPU plumbing path: plain autograd (`.grad`), like every other op here
    return _AddBias.apply(x, bias)


class _MainGradSink(torch.autograd.Function):
    """Identity on a parameter whose gradient, computed by plain autograd ops, belongs in the flat gradient buffer
    (``param.main_grad``, overwrite / accumulate like the kernels' weight gradients) instead of ``param.grad`` --
    the fused optimizer only reads the flat buffer."""

    @staticmethod
    def forward(ctx, param):
        ctx.param = param
        return param.view_as(param)

    @staticmethod
    def backward(ctx, g):
        p = ctx.param
        mg = getattr(p, "main_grad", None)
        if mg is None:
            return g
        if getattr(p, "grad_accumulate", False):
            mg.add_(g.to(mg.dtype))
        else:
            mg.copy_(g)
        return None


def qk_norm(qkv: torch.Tensor, q_weight: torch.Tensor, k_weight: torch.Tensor, n_heads: int, n_kv: int, head_dim: int,
            eps: float) -> torch.Tensor:
    """Qwen3's RMSNorm over ``head_dim`` of every q and k head of the fused ``[T, (Hq + 2 Hkv) D]`` activation (before
    RoPE).  No sm_100a kernel yet: torch ops on every device, counted as a library fallback on the GPU."""
    T = qkv.shape[0]
    qd, kd = n_heads * head_dim, n_kv * head_dim
    if qkv.is_cuda:
        _note_fallback("qk_norm", "per-head RMSNorm of q / k (qwen3) runs as torch ops")
        q_weight, k_weight = _MainGradSink.apply(q_weight), _MainGradSink.apply(k_weight)
    q = _rmsnorm_ref(qkv[:, :qd].reshape(T, n_heads, head_dim), q_weight, eps).reshape(T, qd)
    k = _rmsnorm_ref(qkv[:, qd:qd + kd].reshape(T, n_kv, head_dim), k_weight, eps).reshape(T, kd)
    return torch.cat([q, k, qkv[:, qd + kd:]], 1)


def rope_qkv(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq_len: int,
             n_rot_heads: int, head_dim: int, grad_inplace: bool = False) -> torch.Tensor:
    """Rotate the q and k heads of a fused ``[tokens, (Hq+2Hkv)*D]`` activation in place.

    ``grad_inplace=True`` lets backward rotate the incoming gradient in place as well; only
    safe when the caller knows that gradient has a single consumer (true for the dqkv tensor
    our attention backward allocates)."""
    if not qkv.is_cuda:
        return _rope_ref(qkv, cos, sin, seq_len, n_rot_heads, head_dim)
    return _RoPE.apply(qkv, cos, sin, seq_len, n_rot_heads, head_dim, grad_inplace)


# -----------------------------------------------------------------------------
# SwiGLU
# -----------------------------------------------------------------------------
class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        rows, two_i = gu.shape
        out = torch.empty((rows, two_i // 2), dtype=gu.dtype, device=gu.device)
        load_C().swiglu_fwd(gu.data_ptr(), out.data_ptr(), rows, two_i // 2, cuda_stream())
        ctx.save_for_backward(gu)
        return out

    @staticmethod
    def backward(ctx, dout):
        (gu,) = ctx.saved_tensors
        rows, two_i = gu.shape
        dgu = torch.empty_like(gu)
        load_C().swiglu_bwd(dout.contiguous().data_ptr(), gu.data_ptr(), dgu.data_ptr(), rows,
                            two_i // 2, cuda_stream())
        return dgu


def swiglu(gu: torch.Tensor) -> torch.Tensor:
    """``silu(gate) * up`` on a fused ``[rows, 2*I]`` activation whose columns INTERLEAVE gate and up
    (``gu[:, 2i] = gate_i``, ``gu[:, 2i+1] = up_i``; the rows of the fused gate_up weight are interleaved the
    same way -- ``interleave_gate_up`` -- so that a 256-wide GEMM tile holds whole pairs and the GEMM epilogue
    can apply the activation itself)."""
    if not gu.is_cuda:
        g, u = gu[..., 0::2], gu[..., 1::2]
        return (F.silu(g.float()) * u.float()).to(gu.dtype)
    assert gu.dim() == 2 and gu.is_contiguous()
    return _SwiGLU.apply(gu)


# -----------------------------------------------------------------------------
# Attention (causal, GQA) on the fused qkv activation
# -----------------------------------------------------------------------------
def _attention_ref(qkv, batch, seq, n_heads, n_kv, head_dim):
    T = batch * seq
    q = qkv[:, : n_heads * head_dim].view(batch, seq, n_heads, head_dim).transpose(1, 2)
    k = qkv[:, n_heads * head_dim: (n_heads + n_kv) * head_dim].view(batch, seq, n_kv, head_dim).transpose(1, 2)
    v = qkv[:, (n_heads + n_kv) * head_dim:].view(batch, seq, n_kv, head_dim).transpose(1, 2)
    rep = n_heads // n_kv
    k = k.repeat_interleave(rep, dim=1)
    v = v.repeat_interleave(rep, dim=1)
    o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
    return o.transpose(1, 2).reshape(T, n_heads * head_dim)


class _FlashAttnQKV(torch.autograd.Function):
    """Causal GQA attention reading q/k/v as strided views of the fused qkv buffer and
    writing dq/dk/dv straight into one dqkv buffer (no concat / split copies)."""

    @staticmethod
    def forward(ctx, qkv, batch, seq, n_heads, n_kv, head_dim):
        from flash_attn.flash_attn_interface import _flash_attn_forward

        q = qkv[:, : n_heads * head_dim].view(batch, seq, n_heads, head_dim)
        k = qkv[:, n_heads * head_dim: (n_heads + n_kv) * head_dim].view(batch, seq, n_kv, head_dim)
        v = qkv[:, (n_heads + n_kv) * head_dim:].view(batch, seq, n_kv, head_dim)
        scale = 1.0 / math.sqrt(head_dim)
        out, lse, _, rng = _flash_attn_forward(q, k, v, 0.0, scale, True, -1, -1, 0.0, None, False)
        ctx.save_for_backward(qkv, out, lse, rng)
        ctx.dims = (batch, seq, n_heads, n_kv, head_dim, scale)
        return out.view(batch * seq, n_heads * head_dim)

    @staticmethod
    def backward(ctx, dout):
        from flash_attn.flash_attn_interface import _flash_attn_backward

        qkv, out, lse, rng = ctx.saved_tensors
        batch, seq, n_heads, n_kv, head_dim, scale = ctx.dims
        q = qkv[:, : n_heads * head_dim].view(batch, seq, n_heads, head_dim)
        k = qkv[:, n_heads * head_dim: (n_heads + n_kv) * head_dim].view(batch, seq, n_kv, head_dim)
        v = qkv[:, (n_heads + n_kv) * head_dim:].view(batch, seq, n_kv, head_dim)
        dqkv = torch.empty_like(qkv)
        dq = dqkv[:, : n_heads * head_dim].view(batch, seq, n_heads, head_dim)
        dk = dqkv[:, n_heads * head_dim: (n_heads + n_kv) * head_dim].view(batch, seq, n_kv, head_dim)
        dv = dqkv[:, (n_heads + n_kv) * head_dim:].view(batch, seq, n_kv, head_dim)
        do = dout.contiguous().view(batch, seq, n_heads, head_dim)
        _flash_attn_backward(do, q, k, v, out, lse, dq, dk, dv, 0.0, scale, True, -1, -1, 0.0, None,
                             False, rng)
        return dqkv, None, None, None, None, None


class _TcgenAttnQKV(torch.autograd.Function):
    """Hand-written tcgen05 flash attention, forward AND backward (csrc/kernels/attention_tcgen05.cu,
    attention_bwd_tcgen05.cu).  Reads q/k/v straight out of the fused qkv activation through TMA and
    writes dq/dk/dv straight into the fused d(qkv) the qkv projection's backward consumes.
    ``HYPHA_ATTN_BWD=flash2`` routes the backward through flash-attn 2 (library) for A/B runs."""

    @staticmethod
    def forward(ctx, qkv, batch, seq, n_heads, n_kv, head_dim, scale=None):
        scale = 1.0 / math.sqrt(head_dim) if scale is None else float(scale)
        out = torch.empty((batch * seq, n_heads * head_dim), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((batch, n_heads, seq), dtype=torch.float32, device=qkv.device)
        load_C().attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), batch, seq, n_heads, n_kv, head_dim,
                               qkv.stride(0), out.stride(0), scale, cuda_stream())
        ctx.save_for_backward(qkv, out, lse)
        ctx.dims = (batch, seq, n_heads, n_kv, head_dim, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        batch, seq, n_heads, n_kv, head_dim, scale = ctx.dims
        dqkv = torch.empty_like(qkv)
        if os.environ.get("HYPHA_ATTN_BWD", "fused") in ("fused", "split", "tcgen05"):
            dout = dout if dout.stride(-1) == 1 and dout.stride(0) % 8 == 0 else dout.contiguous()
            variant = os.environ.get("HYPHA_ATTN_BWD", "fused")
            return attention_bwd_tcgen05(qkv, out, dout, lse, batch, seq, n_heads, n_kv, head_dim, dqkv,
                                         "split" if variant == "tcgen05" else variant, scale=scale), None, None, None, None, None, None
        from flash_attn.flash_attn_interface import _flash_attn_backward

        sl = lambda t, a, b, hh: t[:, a * head_dim: b * head_dim].view(batch, seq, hh, head_dim)
        q, k, v = sl(qkv, 0, n_heads, n_heads), sl(qkv, n_heads, n_heads + n_kv, n_kv), sl(qkv, n_heads + n_kv, n_heads + 2 * n_kv, n_kv)
        dq, dk, dv = sl(dqkv, 0, n_heads, n_heads), sl(dqkv, n_heads, n_heads + n_kv, n_kv), sl(dqkv, n_heads + n_kv, n_heads + 2 * n_kv, n_kv)
        do = dout.contiguous().view(batch, seq, n_heads, head_dim)
        rng = torch.zeros(2, dtype=torch.int64, device=qkv.device)
        _flash_attn_backward(do, q, k, v, out.view(batch, seq, n_heads, head_dim), lse, dq, dk, dv, 0.0, scale, True,
                             -1, -1, 0.0, None, False, rng)
        return dqkv, None, None, None, None, None, None


_ATTN_BWD_WS: dict = {}


def attention_bwd_tcgen05(qkv, out, dout, lse, batch: int, seq: int, n_heads: int, n_kv: int, head_dim: int,
                          dqkv: Optional[torch.Tensor] = None, variant: Optional[str] = None,
                          scale: Optional[float] = None) -> torch.Tensor:
    """Raw backward: returns ``dqkv [tokens, (Hq+2Hkv)*D]`` given the forward's (qkv, out, lse) and dout.

    ``variant`` (default ``HYPHA_ATTN_BWD`` or ``fused``): ``fused`` = one kernel, 5 GEMMs per tile pair, dQ
    partials summed in fp32 by TMA reduce-add (csrc/kernels/attention_bwd_fused.cu); ``split`` = the
    atomics-free dK/dV + dQ kernel pair that recomputes S and dP (7 GEMMs, bit-reproducible;
    csrc/kernels/attention_bwd_tcgen05.cu)."""
    variant = variant or os.environ.get("HYPHA_ATTN_BWD", "fused")
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    C = load_C()
    args = (batch, seq, n_heads, n_kv, head_dim, qkv.stride(0), out.stride(0), dout.stride(0), dqkv.stride(0),
            1.0 / math.sqrt(head_dim) if scale is None else float(scale), cuda_stream())
    if variant == "fused":
        # the workspace (D, lse2 and the fp32 dQ accumulator: 4 bytes per q element) is cached per shape and
        # stream-ordered, so a 32-layer backward reuses one allocation instead of 32 x 134 MB
        n = int(C.attention_bwd_fused_ws_floats(batch, seq, n_heads))
        key = (qkv.device.index, n)
        ws = _ATTN_BWD_WS.get(key)
        if ws is None:
            _ATTN_BWD_WS.clear()
            ws = _ATTN_BWD_WS[key] = torch.empty(n, dtype=torch.float32, device=qkv.device)
        C.attention_bwd_fused(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                              ws.data_ptr(), *args)
    elif variant == "split":
        ws = torch.empty(2 * batch * n_heads * seq, dtype=torch.float32, device=qkv.device)
        C.attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                        ws.data_ptr(), *args)
    else:
        raise ValueError(f"unknown attention backward variant {variant!r}")
    return dqkv


def set_attention_fwd_version(version: int) -> None:
    """2 (default) = two 128-row Q tiles per CTA ping-ponging on one MMA issuer, 128-key tiles (needs
    seq % 256 == 0, else version 1 runs); 3 = the same kernel made persistent with a separate epilogue
    warpgroup; 4 = two softmax warpgroups per Q tile; 1 = one Q tile per CTA, 64-key tiles, two CTAs per SM."""
    load_C().attention_fwd_set_version(int(version))


def set_attention_fwd_poly(n: int) -> None:
    """How many of every 4 exponential pairs the forward evaluates with the FMA-pipe polynomial instead of
    MUFU.EX2 (0, 1 or 2; the softmax is MUFU-bound on B200)."""
    load_C().attention_fwd_set_poly(int(n))


def attention_fwd_tcgen05(qkv: torch.Tensor, batch: int, seq: int, n_heads: int, n_kv: int, head_dim: int,
                          version: Optional[int] = None):
    """Raw forward: returns ``(out [tokens, Hq*D], lse [B, Hq, S])``."""
    if version is not None:
        prev = load_C().attention_fwd_get_version()
        set_attention_fwd_version(version)
        try:
            return attention_fwd_tcgen05(qkv, batch, seq, n_heads, n_kv, head_dim)
        finally:
            set_attention_fwd_version(prev)
    out = torch.empty((batch * seq, n_heads * head_dim), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((batch, n_heads, seq), dtype=torch.float32, device=qkv.device)
    load_C().attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), batch, seq, n_heads, n_kv, head_dim,
                           qkv.stride(0), out.stride(0), 1.0 / math.sqrt(head_dim), cuda_stream())
    return out, lse


_ATTN_BACKEND = os.environ.get("HYPHA_ATTN", "tcgen05")


def set_attention_backend(name: str) -> None:
    """``tcgen05`` (default): the hand-written flash attention, forward and backward, for head_dim 128 and
    sequence lengths that are multiples of 128 -- other shapes run cuDNN.  ``cudnn`` (cuDNN fused attention
    through torch SDPA; library; +6 % on the Llama-3-8B step) and ``flash2`` (flash-attn 2 mma.sync kernels;
    library) are kept for A/B runs."""
    global _ATTN_BACKEND
    if name not in ("cudnn", "flash2", "tcgen05"):
        raise ValueError(name)
    _ATTN_BACKEND = name


def get_attention_backend() -> str:
    return _ATTN_BACKEND


def _sdpa_views(qkv, batch, seq, n_heads, n_kv, head_dim):
    q = qkv[:, : n_heads * head_dim].view(batch, seq, n_heads, head_dim).transpose(1, 2)
    k = qkv[:, n_heads * head_dim: (n_heads + n_kv) * head_dim].view(batch, seq, n_kv, head_dim).transpose(1, 2)
    v = qkv[:, (n_heads + n_kv) * head_dim:].view(batch, seq, n_kv, head_dim).transpose(1, 2)
    return q, k, v


def _sdpa_call(q, k, v, gqa: bool):
    from torch.nn.attention import SDPBackend, sdpa_kernel

    with sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION]):
        return F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=gqa)


class _CudnnAttnQKV(torch.autograd.Function):
    """cuDNN fused attention (library) on views of the fused qkv activation.  The inner SDPA graph is
    kept private so backward receives dq/dk/dv as three tensors and assembles ``dqkv`` with ONE
    concatenation instead of autograd's three zero-fills + three slice copies + two adds."""

    @staticmethod
    def forward(ctx, qkv, batch, seq, n_heads, n_kv, head_dim):
        with torch.enable_grad():
            base = qkv.detach()
            q, k, v = (t.requires_grad_() for t in _sdpa_views(base, batch, seq, n_heads, n_kv, head_dim))
            o = _sdpa_call(q, k, v, n_kv != n_heads)
        ctx.inner = (q, k, v, o)
        ctx.dims = (batch, seq, n_heads, n_kv, head_dim)
        return o.detach().transpose(1, 2).reshape(batch * seq, n_heads * head_dim)

    @staticmethod
    def backward(ctx, dout):
        q, k, v, o = ctx.inner
        ctx.inner = None
        batch, seq, n_heads, n_kv, head_dim = ctx.dims
        do = dout.view(batch, seq, n_heads, head_dim).transpose(1, 2)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), do)
        T = batch * seq
        dqkv = torch.cat([dq.transpose(1, 2).reshape(T, -1), dk.transpose(1, 2).reshape(T, -1),
                          dv.transpose(1, 2).reshape(T, -1)], dim=1)
        return dqkv, None, None, None, None, None


def _attention_sdpa(qkv, batch, seq, n_heads, n_kv, head_dim):
    if torch.is_grad_enabled() and qkv.requires_grad:
        return _CudnnAttnQKV.apply(qkv, batch, seq, n_heads, n_kv, head_dim)
    q, k, v = _sdpa_views(qkv, batch, seq, n_heads, n_kv, head_dim)
    o = _sdpa_call(q, k, v, n_kv != n_heads)
    return o.transpose(1, 2).reshape(batch * seq, n_heads * head_dim)


def _attention_tcgen05_padded(qkv: torch.Tensor, batch: int, seq: int, n_heads: int, n_kv: int,
                              head_dim: int) -> torch.Tensor:
    """Shapes the tcgen05 kernels are not specialised for -- head_dim < 128 (GPT-2, llama3-1b: 64) and sequences
    that are not a multiple of 128 -- run on them all the same, zero-padded to the kernels' tile: head_dim -> 128
    (zero q/k columns add nothing to a score, zero v columns give zero outputs that are sliced off; the softmax
    scale stays 1/sqrt(head_dim)) and seq -> the next multiple of 128 (under the causal mask no real query sees a
    padded key; padded queries are discarded and receive zero gradient, so they contribute nothing to dK / dV).
    Exact, at the price of the padding's FLOPs and two copies -- no library kernel on the default path."""
    H = n_heads + 2 * n_kv
    S2 = (seq + 127) // 128 * 128
    x = qkv.reshape(batch, seq, H, head_dim)
    xp = torch.zeros((batch, S2, H, 128), dtype=qkv.dtype, device=qkv.device)
    xp[:, :seq, :, :head_dim] = x
    out = _TcgenAttnQKV.apply(xp.view(batch * S2, H * 128), batch, S2, n_heads, n_kv, 128, 1.0 / math.sqrt(head_dim))
    return out.view(batch, S2, n_heads, 128)[:, :seq, :, :head_dim].reshape(batch * seq, n_heads * head_dim)


def attention_qkv(qkv: torch.Tensor, batch: int, seq: int, n_heads: int, n_kv: int,
                  head_dim: int) -> torch.Tensor:
    """Causal self-attention on a fused (already RoPE'd) qkv activation -> ``[tokens, Hq*D]``."""
    if not qkv.is_cuda:
        return _attention_ref(qkv, batch, seq, n_heads, n_kv, head_dim)
    if (_ATTN_BACKEND == "tcgen05" and head_dim == 128 and seq % 128 == 0 and qkv.dtype == torch.bfloat16
            and n_heads % n_kv == 0 and qkv.stride(-1) == 1 and qkv.stride(0) % 8 == 0):
        return _TcgenAttnQKV.apply(qkv, batch, seq, n_heads, n_kv, head_dim)
    if (_ATTN_BACKEND == "tcgen05" and qkv.dtype == torch.bfloat16 and n_heads % n_kv == 0 and 0 < head_dim <= 128
            and head_dim % 8 == 0):
        return _attention_tcgen05_padded(qkv, batch, seq, n_heads, n_kv, head_dim)
    if _ATTN_BACKEND == "flash2":
        return _FlashAttnQKV.apply(qkv, batch, seq, n_heads, n_kv, head_dim)
    if _ATTN_BACKEND == "tcgen05":
        _note_fallback("attention", f"tcgen05 kernels cover bf16, head_dim <= 128 (multiple of 8) and an integral GQA ratio "
                                    f"(got head_dim={head_dim}, {qkv.dtype}, heads {n_heads}/{n_kv}): cuDNN used")
    return _attention_sdpa(qkv, batch, seq, n_heads, n_kv, head_dim)      # cudnn, or a shape tcgen05 does not cover


# -----------------------------------------------------------------------------
# Embedding
# -----------------------------------------------------------------------------
class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table):
        tokens = ids.numel()
        dim = table.shape[1]
        out = torch.empty((tokens, dim), dtype=table.dtype, device=table.device)
        load_C().embedding_fwd(ids.data_ptr(), table.data_ptr(), out.data_ptr(), tokens, dim,
                               cuda_stream())
        ctx.save_for_backward(ids)
        ctx.table = table
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        table = ctx.table
        mg = getattr(table, "main_grad", None)
        if mg is not None:
            # a table that doubles as the LM head (`tied_head`) already holds the head's weight gradient here
            if not (getattr(table, "grad_accumulate", False) or getattr(table, "tied_head", False)):
                mg.zero_()
            target, ret = mg, None
        else:
            target = torch.zeros_like(table)
            ret = target
        load_C().embedding_bwd(ids.data_ptr(), dout.contiguous().data_ptr(), target.data_ptr(),
                               ids.numel(), table.shape[1], cuda_stream())
        return None, ret


def embedding(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """Row gather ``table[ids]`` -> ``[tokens, dim]`` (ids flattened, int64)."""
    if not table.is_cuda:
        return F.embedding(ids.reshape(-1), table)
    return _Embedding.apply(ids.reshape(-1).contiguous(), table)


# -----------------------------------------------------------------------------
# Fused LM head + softmax cross-entropy (chunked; logits are never kept)
# -----------------------------------------------------------------------------
class _LinearCrossEntropy(torch.autograd.Function):
    """loss = mean CE(h @ W^T, targets).  The backward of logits is produced in place by the
    CE kernel inside forward, immediately consumed by the dgrad/wgrad GEMMs, and dropped, so
    peak memory is one ``[chunk, vocab]`` bf16 tile instead of ``[tokens, vocab]`` fp32."""

    @staticmethod
    def forward(ctx, h, weight, targets, chunk, ignore_index, n_valid=None):
        C = load_C()
        T, K = h.shape
        V = weight.shape[0]
        if ignore_index is None:
            n_valid, ignore_index = T, -100          # caller guarantees every target is valid
        elif n_valid is not None:
            n_valid = max(int(n_valid), 1)           # the caller knows how many targets it masked: no host sync
        else:
            n_valid = max(int((targets != ignore_index).sum().item()), 1)   # host sync
        scale = 1.0 / n_valid
        loss_rows = torch.empty(T, dtype=torch.float32, device=h.device)
        dh = torch.empty_like(h)
        logits = torch.empty((min(chunk, T), V), dtype=torch.bfloat16, device=h.device)
        mg = getattr(weight, "main_grad", None)
        dw = mg if mg is not None else torch.empty_like(weight)
        accumulate = bool(getattr(weight, "grad_accumulate", False)) if mg is not None else False
        for s in range(0, T, chunk):
            e = min(s + chunk, T)
            lg = logits[: e - s]
            gemm(h[s:e], weight, out=lg)
            C.cross_entropy_fwd_bwd(lg.data_ptr(), targets[s:e].data_ptr(), loss_rows[s:e].data_ptr(),
                                    e - s, V, lg.stride(0), scale, ignore_index, cuda_stream())
            gemm(lg, weight, a_kmajor=True, b_kmajor=False, out=dh[s:e])
            gemm(lg, h[s:e], a_kmajor=False, b_kmajor=False, out=dw, accumulate=accumulate or s > 0)
        ctx.save_for_backward(dh)
        ctx.dw = None if mg is not None else dw
        return loss_rows.sum() * scale

    @staticmethod
    def backward(ctx, dloss):
        (dh,) = ctx.saved_tensors
        # dloss is 1.0 for a plain `loss.backward()`; keep generality without a host sync.
        dh = dh * dloss.to(dh.dtype)
        dw = ctx.dw
        if dw is not None:
            dw = dw * dloss.to(dw.dtype)
        return dh, dw, None, None, None, None


def linear_cross_entropy(h: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor,
                         chunk: int = 4096, ignore_index: Optional[int] = None,
                         n_valid: Optional[int] = None) -> torch.Tensor:
    """Mean token cross-entropy of ``h @ weight.T`` against ``targets`` (LM head fused in).

    ``ignore_index=None`` promises that every target is valid, which keeps the op free of
    host synchronisation; pass an index (e.g. ``-100``) to mask padding tokens -- and ``n_valid`` (how many targets
    are NOT masked) when the caller knows it, which keeps the masked variant free of the synchronisation as well."""
    if not h.is_cuda:
        logits = F.linear(h, weight).float()
        return F.cross_entropy(logits, targets.reshape(-1),
                               ignore_index=-100 if ignore_index is None else ignore_index)
    return _LinearCrossEntropy.apply(h, weight, targets.reshape(-1).contiguous(), chunk, ignore_index, n_valid)


def cross_entropy_inplace(logits: torch.Tensor, targets: torch.Tensor, grad_scale: float,
                          ignore_index: int = -100) -> torch.Tensor:
    """Raw kernel access: returns per-row losses and overwrites ``logits`` with d(loss)/d(logits)."""
    rows, vocab = logits.shape
    loss_rows = torch.empty(rows, dtype=torch.float32, device=logits.device)
    load_C().cross_entropy_fwd_bwd(logits.data_ptr(), targets.data_ptr(), loss_rows.data_ptr(), rows,
                                   vocab, logits.stride(0), grad_scale, ignore_index, cuda_stream())
    return loss_rows
