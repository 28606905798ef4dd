#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Installs the UNMODIFIED reference (sslap v0.2.5, Cython) into oracle/_ref/
# so tests / bench.py's cpu_baseline + `--impl reference` arm can run it side by side with the CUDA path.
# The reference tree is read-only and its setup.py builds in-tree, so we build from a scratch copy in /tmp;
# nothing from the reference's sources is ever written into tracked files (oracle/_ref/ is git-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${1:-/root/reference}"
if [ ! -d "$REF/sslap" ]; then echo "reference tree not found at $REF (GPU box uses the prebuilt oracle/_ref)"; exit 0; fi
TMP="$(mktemp -d /tmp/sslap_src.XXXXXX)"
cp -r "$REF/." "$TMP/"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" >/dev/null
rm -rf "$TMP"
echo "reference installed into $HERE/_ref"
