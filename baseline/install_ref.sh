#!/usr/bin/env bash
# Installs the UNMODIFIED reference (scart97/thunder-speech 3.2.0) into baseline/_ref/ (git-ignored, NOT gpurun-ignored,
# so it travels to the GPU box) for `bench.py --impl reference` / `--impl reference_cuda`.
#
#   1. the contract's command: pip install --no-index --no-build-isolation --target baseline/_ref <copy of /root/reference>
#      -> fails in this image: the project's build backend is poetry-core ("No module named 'poetry'"), absent offline.
#   2. fallback = exactly what that install would have put there with --no-deps: the package directory `thunder/`
#      (pyproject: packages = [{include = "thunder", from = "src"}]) copied byte for byte, plus a minimal
#      thunder_speech-3.2.0.dist-info so that `thunder/__init__.py`'s importlib.metadata.version("thunder-speech") resolves.
# Nothing under baseline/_ref is tracked by git; nothing in the product or the tests imports it.
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${THUNDER_REF:-/root/reference}"
DST="$HERE/_ref"
[ -d "$REF/src/thunder" ] || { echo "no reference at $REF (GPU box?): keeping whatever is in $DST"; exit 0; }
rm -rf "$DST" /tmp/_thunder_ref_copy
cp -r "$REF" /tmp/_thunder_ref_copy
if python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$DST" \
      /tmp/_thunder_ref_copy >/tmp/_thunder_ref_pip.log 2>&1; then
  echo "pip install ok -> $DST"
else
  echo "pip install failed ($(grep -m1 -o "No module named '[a-z_]*'" /tmp/_thunder_ref_pip.log)): copying the package directory instead"
  mkdir -p "$DST/thunder_speech-3.2.0.dist-info"
  cp -r "$REF/src/thunder" "$DST/thunder"
  printf 'Metadata-Version: 2.1\nName: thunder-speech\nVersion: 3.2.0\n' > "$DST/thunder_speech-3.2.0.dist-info/METADATA"
  printf 'manual copy (poetry-core unavailable offline)\n' > "$DST/thunder_speech-3.2.0.dist-info/INSTALLER"
fi
find "$DST" -name __pycache__ -prune -exec rm -rf {} +
( cd "$REF/src" && find thunder -name '*.py' | sort | xargs sha256sum ) > /tmp/_ref_a.sha
( cd "$DST" && find thunder -name '*.py' | sort | xargs sha256sum ) > /tmp/_ref_b.sha
cmp -s /tmp/_ref_a.sha /tmp/_ref_b.sha && echo "baseline/_ref/thunder is byte-identical to $REF/src/thunder ($(wc -l < /tmp/_ref_a.sha) files)"
rm -rf /tmp/_thunder_ref_copy
