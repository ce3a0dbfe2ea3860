#!/usr/bin/env python
"""Build a BINARY release artefact (the reference ships prebuilt role binaries, install.sh:1-601 +
.github/workflows/release.yaml:105-179; this is the counterpart for a Python + native-extension tree).

    python tools/make_release.py [--out dist] [--skip-cuda]

Produces ``dist/hypha-b200-<version>-<pytag>-linux_x86_64-sm100a.tar.gz`` holding the Python package WITH the
prebuilt ``_core`` / ``_C`` shared objects (no compiler, no nvcc and no source build on the target box), the role
launchers, ``install.sh``, the docs, and a ``MANIFEST.json`` with the sha256 of every file.  Install it with

    HYPHA_RELEASE=dist/hypha-b200-....tar.gz sh install.sh            # or: sh install.sh --from-release <tar.gz | URL>

which verifies the manifest, unpacks under ``$HYPHA_PREFIX`` (default ``~/.local/share/hypha-b200``) and links the
``hypha-*`` commands; nothing is compiled."""
import argparse
import hashlib
import io
import json
import os
import sys
import sysconfig
import tarfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

INCLUDE_DIRS = ["hypha_b200", "bin", "docs", "rfc"]
INCLUDE_FILES = ["install.sh", "README.md", "CHANGELOG.md", "pyproject.toml", "bench.py", "__graft_entry__.py"]
SKIP_DIR_NAMES = {"__pycache__", "csrc"}            # sources of the native code are not needed to RUN a release


def version() -> str:
    import re

    txt = open(os.path.join(ROOT, "pyproject.toml")).read()
    m = re.search(r'^version\s*=\s*"([^"]+)"', txt, re.M)
    return m.group(1) if m else "0.0.0"


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "dist"))
    ap.add_argument("--skip-cuda", action="store_true", help="CPU-only artefact (control plane + C++ core)")
    a = ap.parse_args()
    from hypha_b200.build import build_core, build_cuda

    build_core(False)
    if not a.skip_cuda:
        build_cuda(False)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    need = ["hypha_b200/_core" + ext] + ([] if a.skip_cuda else ["hypha_b200/_C" + ext])
    for n in need:
        if not os.path.exists(os.path.join(ROOT, n)):
            print(f"missing built extension {n}", file=sys.stderr)
            return 1
    pytag = f"cp{sys.version_info.major}{sys.version_info.minor}"
    arch = "cpu" if a.skip_cuda else "sm100a"
    name = f"hypha-b200-{version()}-{pytag}-linux_x86_64-{arch}"
    os.makedirs(a.out, exist_ok=True)
    files = []
    for d in INCLUDE_DIRS:
        for base, dirs, fs in os.walk(os.path.join(ROOT, d)):
            dirs[:] = [x for x in dirs if x not in SKIP_DIR_NAMES]
            for f in fs:
                if f.endswith((".pyc", ".o")) or (f.endswith(".so") and a.skip_cuda and f.startswith("_C")):
                    continue
                files.append(os.path.relpath(os.path.join(base, f), ROOT))
    files += [f for f in INCLUDE_FILES if os.path.exists(os.path.join(ROOT, f))]
    manifest = {"name": name, "version": version(), "python": pytag, "arch": arch,
                "built": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "files": {}}
    for f in sorted(files):
        manifest["files"][f] = hashlib.sha256(open(os.path.join(ROOT, f), "rb").read()).hexdigest()
    out = os.path.join(a.out, name + ".tar.gz")
    with tarfile.open(out, "w:gz") as tar:
        for f in sorted(files):
            tar.add(os.path.join(ROOT, f), arcname=f"{name}/{f}")
        blob = json.dumps(manifest, indent=1).encode()
        info = tarfile.TarInfo(f"{name}/MANIFEST.json")
        info.size = len(blob)
        info.mtime = int(time.time())
        tar.addfile(info, io.BytesIO(blob))
    digest = hashlib.sha256(open(out, "rb").read()).hexdigest()
    open(out + ".sha256", "w").write(f"{digest}  {os.path.basename(out)}\n")
    print(json.dumps({"artefact": out, "sha256": digest, "files": len(files), "extensions": need}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
