"""In-tree native build for hypha_b200.

Two shared objects are produced next to this file:

* ``_core``  -- pure C++17 control-plane core (ledger, resources, offer
  aggregation, simulation, trackers, windowing).  No CUDA dependency, so the
  CPU test-suite can always import it.
* ``_C``     -- sm_100a CUDA kernels + peer-memory transport.  Every ``.cu`` is
  compiled by nvcc with ``-gencode arch=compute_100a,code=sm_100a -lineinfo``;
  the binding layer takes raw device pointers / stream handles so no torch C++
  headers are needed (a full rebuild is seconds, not minutes).

Objects are cached under ``build/`` keyed by a hash of the source, its local
includes and the flags.  ``python -m hypha_b200.build`` builds everything.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = ROOT / "build"
EXT = sysconfig.get_config_var("EXT_SUFFIX") or ".so"

CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-Xcudafe", "--diag_suppress=177", "-DNDEBUG",
]
CXX = os.environ.get("CXX", "g++")
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-fvisibility=hidden", "-DNDEBUG"]


def _py_includes() -> list[str]:
    import pybind11

    return ["-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"]]


_INC_RE = re.compile(r'^\s*#\s*include\s*"([^"]+)"', re.M)


def _source_hash(src: Path, flags: list[str], seen: set[Path] | None = None) -> str:
    seen = seen if seen is not None else set()
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())

    def visit(p: Path) -> None:
        if p in seen or not p.exists():
            return
        seen.add(p)
        text = p.read_text()
        h.update(str(p.relative_to(ROOT)).encode())
        h.update(text.encode())
        for inc in _INC_RE.findall(text):
            for base in (p.parent, CSRC, CSRC / "kernels", CSRC / "core", CSRC / "comm"):
                q = (base / inc).resolve()
                if q.exists():
                    visit(q)
                    break

    visit(src)
    return h.hexdigest()[:16]


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build command failed: {' '.join(cmd[:6])} ...")
    if verbose and (r.stdout or r.stderr):
        sys.stderr.write(r.stdout + r.stderr)


def _flags(kind: str, extra: list[str] | None = None) -> list[str]:
    if kind == "cu":
        return NVCC_FLAGS + ["-I", str(CSRC), "-I", str(CSRC / "kernels")] + (extra or [])
    return CXX_FLAGS + ["-I", str(CSRC), "-I", os.path.join(CUDA_HOME, "include")] + _py_includes() + (extra or [])


def _up_to_date(out: Path, parts: list[tuple[Path, str]]) -> tuple[bool, Path, str]:
    """Is the shared object next to the package the build of the current sources?  Its key (the hashes of every
    source with its flags) is kept IN the package directory, so a copy of the tree that carries the `.so` but not the
    object cache under ``build/`` (a fresh GPU box) does not compile anything."""
    stamp = out.with_suffix(".srckey")
    key = ",".join(f"{src.stem}.{_source_hash(src, _flags(kind))}" for src, kind in parts)
    try:
        return out.exists() and stamp.read_text() == key, stamp, key
    except OSError:
        return False, stamp, key


def _compile(src: Path, kind: str, verbose: bool, extra: list[str] | None = None) -> Path:
    flags = _flags(kind, extra)
    tool = [NVCC] if kind == "cu" else [CXX]
    tag = _source_hash(src, flags)
    obj = BUILD / f"{src.stem}.{tag}.o"
    if obj.exists():
        return obj
    BUILD.mkdir(parents=True, exist_ok=True)
    for stale in BUILD.glob(f"{src.stem}.*.o"):
        stale.unlink()
    _run(tool + flags + ["-c", str(src), "-o", str(obj)], verbose)
    return obj


def cuda_sources() -> list[Path]:
    return sorted((CSRC / "kernels").glob("*.cu"))


def build_core(verbose: bool = False) -> Path:
    """Build the pure-C++ control-plane core (``hypha_b200/_core*.so``)."""
    out = PKG / f"_core{EXT}"
    fresh, srckey, key_src = _up_to_date(out, [(CSRC / "core" / "core_bindings.cpp", "cpp")])
    if fresh:
        return out
    obj = _compile(CSRC / "core" / "core_bindings.cpp", "cpp", verbose)
    stamp = BUILD / "_core.stamp"
    if out.exists() and stamp.exists() and stamp.read_text() == obj.name:
        srckey.write_text(key_src)
        return out
    _run([CXX, "-shared", "-o", str(out), str(obj), "-pthread"], verbose)
    stamp.write_text(obj.name)
    srckey.write_text(key_src)
    return out


def build_cuda(verbose: bool = False) -> Path:
    """Build the sm_100a kernel + peer-memory extension (``hypha_b200/_C*.so``)."""
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}")
    out = PKG / f"_C{EXT}"
    srcs = cuda_sources()
    cpp = [CSRC / "comm" / "peer_mem.cpp", CSRC / "comm" / "vmm.cpp", CSRC / "bindings.cpp"]
    fresh, srckey, key_src = _up_to_date(out, [(s, "cu") for s in srcs] + [(c, "cpp") for c in cpp])
    if fresh:
        return out
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, "cu", verbose), srcs))
    objs.append(_compile(CSRC / "comm" / "peer_mem.cpp", "cpp", verbose))
    objs.append(_compile(CSRC / "comm" / "vmm.cpp", "cpp", verbose))
    objs.append(_compile(CSRC / "bindings.cpp", "cpp", verbose))
    stamp = BUILD / "_C.stamp"
    key = ",".join(o.name for o in objs)
    if out.exists() and stamp.exists() and stamp.read_text() == key:
        srckey.write_text(key_src)
        return out
    _run([NVCC] + ARCH + ["-shared", "-o", str(out)] + [str(o) for o in objs]
         + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"], verbose)
    stamp.write_text(key)
    srckey.write_text(key_src)
    return out


def build_all(verbose: bool = False) -> None:
    build_core(verbose)
    build_cuda(verbose)


def clean() -> None:
    shutil.rmtree(BUILD, ignore_errors=True)
    for p in PKG.glob("_C*.so"):
        p.unlink()
    for p in PKG.glob("_core*.so"):
        p.unlink()


if __name__ == "__main__":
    if "clean" in sys.argv[1:]:
        clean()
    else:
        build_all(verbose="-v" in sys.argv[1:])
        print("built:", *(p.name for p in sorted(PKG.glob("_*.so"))))
