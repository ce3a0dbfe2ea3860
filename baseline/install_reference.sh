#!/bin/sh
# The one allowed offline install of the reference (outcome recorded in DESIGN.md section 5).
# /root/reference itself is a Rust workspace without pyproject.toml; its only Python package is
# executors/accelerate, whose build backend (uv_build) is not in /opt/wheelhouse.  The package is
# therefore built from a /tmp copy whose [build-system] table points at setuptools instead; the
# Python sources are installed byte-for-byte (checked with diff -r below).
set -eu
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${1:-/root/reference/executors/accelerate}
TMP=$(mktemp -d /tmp/hypha_ref_XXXXXX)
cp -r "$SRC"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'PY'
import sys
p = sys.argv[1]
s = open(p).read()
s = s.replace('requires = ["uv_build>=0.9.8,<0.10.0"]\nbuild-backend = "uv_build"',
              'requires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"')
s += '\n[tool.setuptools.packages.find]\nwhere = ["src"]\nnamespaces = true\n'
open(p, "w").write(s)
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP"
diff -r -x __pycache__ "$HERE/_ref/hypha/accelerate_executor" "$SRC/src/hypha/accelerate_executor" \
    && echo "baseline/_ref: reference executor sources installed unmodified"
rm -rf "$TMP"
