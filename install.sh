#!/bin/sh
# hypha-b200 installer (POSIX sh).  Two modes, like the reference's release installer (install.sh:1-601):
#   * from a checkout (default): build the native extensions in place and link the role launchers;
#   * from a BINARY release (tools/make_release.py): `sh install.sh --from-release <tar.gz | URL>` or
#     HYPHA_RELEASE=<tar.gz | URL> -- verifies sha256 (a `<file>.sha256` next to it, or HYPHA_RELEASE_SHA256) and the
#     per-file MANIFEST.json, unpacks under $HYPHA_PREFIX and links the launchers; nothing is compiled.
#
#   HYPHA_PREFIX           where a release is unpacked (default: $HOME/.local/share/hypha-b200)
#   HYPHA_GITHUB_TOKEN     sent as a bearer token when the release is fetched from a URL
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
RELEASE=${HYPHA_RELEASE:-}
if [ "${1:-}" = "--from-release" ]; then
  [ -n "${2:-}" ] || err "--from-release needs a path or URL"
  RELEASE=$2
fi

link_launchers() {   # $1 = directory that holds the hypha_b200 package
  mkdir -p "$DEST"
  for role in gateway scheduler worker data certutil box; do
    target="$DEST/hypha-$role"
    printf '#!/bin/sh\nPYTHONPATH="%s${PYTHONPATH:+:$PYTHONPATH}" exec %s -m hypha_b200.cli.%s "$@"\n' "$1" "$PY" "$role" > "$target"
    chmod +x "$target"
    say "installed $target"
  done
  case ":$PATH:" in
    *":$DEST:"*) ;;
    *) [ "${HYPHA_NO_MODIFY_PATH:-0}" = "1" ] || say "add $DEST to your PATH to use the hypha-* commands" ;;
  esac
}

if [ -n "$RELEASE" ]; then
  command -v "$PY" >/dev/null 2>&1 || err "python interpreter '$PY' not found"
  PREFIX=${HYPHA_PREFIX:-"$HOME/.local/share/hypha-b200"}
  TMP=$(mktemp -d)
  trap 'rm -rf "$TMP"' EXIT
  case "$RELEASE" in
    http://*|https://*)
      say "downloading $RELEASE"
      if command -v curl >/dev/null 2>&1; then
        curl -fsSL ${HYPHA_GITHUB_TOKEN:+-H "Authorization: Bearer $HYPHA_GITHUB_TOKEN"} -o "$TMP/release.tar.gz" "$RELEASE" || err "download failed"
        curl -fsSL ${HYPHA_GITHUB_TOKEN:+-H "Authorization: Bearer $HYPHA_GITHUB_TOKEN"} -o "$TMP/release.tar.gz.sha256" "$RELEASE.sha256" 2>/dev/null || true
      else
        err "curl is needed to fetch a release from a URL"
      fi
      TAR="$TMP/release.tar.gz" ;;
    *) [ -f "$RELEASE" ] || err "no such release file: $RELEASE"
       TAR=$RELEASE
       [ -f "$RELEASE.sha256" ] && cp "$RELEASE.sha256" "$TMP/release.tar.gz.sha256" ;;
  esac
  WANT=${HYPHA_RELEASE_SHA256:-}
  [ -z "$WANT" ] && [ -f "$TMP/release.tar.gz.sha256" ] && WANT=$(cut -d' ' -f1 "$TMP/release.tar.gz.sha256")
  if [ -n "$WANT" ]; then
    GOT=$("$PY" -c 'import hashlib,sys; print(hashlib.sha256(open(sys.argv[1],"rb").read()).hexdigest())' "$TAR")
    [ "$GOT" = "$WANT" ] || err "sha256 mismatch: expected $WANT, got $GOT"
    say "sha256 verified"
  else
    say "warning: no sha256 to verify the artefact against"
  fi
  mkdir -p "$PREFIX"
  tar -xzf "$TAR" -C "$TMP"
  NAME=$(ls "$TMP" | grep '^hypha-b200-' | head -n 1)
  [ -n "$NAME" ] && [ -f "$TMP/$NAME/MANIFEST.json" ] || err "not a hypha-b200 release (no MANIFEST.json)"
  "$PY" - "$TMP/$NAME" <<'PYEOF' || err "release content does not match its manifest"
import hashlib, json, os, sys
root = sys.argv[1]
man = json.load(open(os.path.join(root, "MANIFEST.json")))
bad = [f for f, h in man["files"].items()
       if not os.path.isfile(os.path.join(root, f)) or hashlib.sha256(open(os.path.join(root, f), "rb").read()).hexdigest() != h]
want = "cp%d%d" % sys.version_info[:2]
if man.get("python") != want:
    sys.stderr.write("release was built for %s, this interpreter is %s\n" % (man.get("python"), want)); sys.exit(1)
sys.exit(1 if bad else 0)
PYEOF
  rm -rf "$PREFIX/$NAME"
  mv "$TMP/$NAME" "$PREFIX/$NAME"
  ln -sfn "$PREFIX/$NAME" "$PREFIX/current"
  PYTHONPATH="$PREFIX/current" HYPHA_NO_AUTOBUILD=1 "$PY" -c 'import hypha_b200._core' || err "the prebuilt core does not import with this interpreter"
  say "unpacked $NAME under $PREFIX (prebuilt extensions, no compilation)"
  link_launchers "$PREFIX/current"
  say "done.  try: hypha-gateway init -o gateway.toml"
  exit 0
fi

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

link_launchers "$ROOT"
say "done.  try: hypha-gateway init -o gateway.toml"
