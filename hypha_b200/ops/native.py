"""Loader for the in-tree native extensions.

``_C`` (sm_100a kernels) must be present on a GPU box: if CUDA is available and
the extension cannot be imported we fail loudly instead of silently running a
PyTorch fallback (the round-end driver checks that the ``.so`` was really
loaded).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
from functools import lru_cache

import torch


class NativeExtensionMissing(RuntimeError):
    pass


@lru_cache(maxsize=None)
def load_C():
    try:
        return importlib.import_module("hypha_b200._C")
    except ImportError as first:
        if os.environ.get("HYPHA_NO_AUTOBUILD") == "1":
            raise NativeExtensionMissing(f"hypha_b200._C is not built: {first}") from first
        try:
            from hypha_b200.build import build_cuda

            build_cuda()
            importlib.invalidate_caches()
            return importlib.import_module("hypha_b200._C")
        except Exception as exc:  # pragma: no cover - only on broken toolchains
            raise NativeExtensionMissing(
                f"hypha_b200._C missing and in-tree build failed: {exc}") from exc


@lru_cache(maxsize=None)
def load_core():
    override = os.environ.get("HYPHA_CORE_OVERRIDE")        # e.g. the ASAN build from tools/sanitize.sh
    if override:
        import glob

        path = glob.glob(os.path.join(override, "_core*.so"))[0]
        spec = importlib.util.spec_from_file_location("hypha_b200._core", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    try:
        return importlib.import_module("hypha_b200._core")
    except ImportError as first:
        if os.environ.get("HYPHA_NO_AUTOBUILD") == "1":
            raise NativeExtensionMissing(f"hypha_b200._core is not built: {first}") from first
        from hypha_b200.build import build_core

        build_core()
        importlib.invalidate_caches()
        return importlib.import_module("hypha_b200._core")


def cuda_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load_C().launch_count())


def reset_launch_count() -> None:
    load_C().reset_launch_count()
