#!/bin/sh
# hypha-b200 installer (POSIX sh).  The reference installer downloads release binaries
# (install.sh:1-601); this tree is source-only and there is no network on a training box, so
# the installer builds the native extensions in place and links the role launchers.
#
#   HYPHA_INSTALL_DIR      where to link bin/hypha-* (default: $HOME/.local/bin)
#   HYPHA_NO_MODIFY_PATH   set to 1 to skip the PATH hint
#   HYPHA_PYTHON           python interpreter to use (default: python)
#   HYPHA_SKIP_CUDA        set to 1 to build only the CPU control-plane core
set -eu

say() { printf '%s\n' "hypha-b200: $*"; }
err() { say "error: $*" >&2; exit 1; }

ROOT=$(CDPATH= cd -- "$(dirname -- "$0")" && pwd)
PY=${HYPHA_PYTHON:-python}
DEST=${HYPHA_INSTALL_DIR:-"$HOME/.local/bin"}

command -v "$PY" >/dev/null 2>&1 || err "python interpreter '$PY' not found"
"$PY" -c 'import sys; sys.exit(0 if sys.version_info >= (3, 10) else 1)' || err "python >= 3.10 required"
"$PY" -c 'import torch, pybind11, cbor2, aiohttp' 2>/dev/null || err "missing python deps (torch, pybind11, cbor2, aiohttp)"

cd "$ROOT"
if [ "${HYPHA_SKIP_CUDA:-0}" = "1" ]; then
  say "building the control-plane core (C++)"
  "$PY" -c 'from hypha_b200.build import build_core; build_core(True)'
else
  command -v nvcc >/dev/null 2>&1 || [ -x /usr/local/cuda/bin/nvcc ] || err "nvcc not found (set HYPHA_SKIP_CUDA=1 for a CPU-only install)"
  say "building _core (C++) and _C (sm_100a CUDA)"
  "$PY" -m hypha_b200.build
fi

mkdir -p "$DEST"
for role in gateway scheduler worker data certutil box; do
  target="$DEST/hypha-$role"
  printf '#!/bin/sh\nPYTHONPATH="%s${PYTHONPATH:+:$PYTHONPATH}" exec %s -m hypha_b200.cli.%s "$@"\n' "$ROOT" "$PY" "$role" > "$target"
  chmod +x "$target"
  say "installed $target"
done

case ":$PATH:" in
  *":$DEST:"*) ;;
  *) [ "${HYPHA_NO_MODIFY_PATH:-0}" = "1" ] || say "add $DEST to your PATH to use the hypha-* commands" ;;
esac
say "done.  try: hypha-gateway init -o gateway.toml"
