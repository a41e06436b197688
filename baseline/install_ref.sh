#!/usr/bin/env bash
# Offline install of the UNMODIFIED reference into the git-ignored baseline/_ref/ (travels to the GPU box with gpurun).
#
#   bash baseline/install_ref.sh [/root/reference]
#
# /root/reference is read-only, so the build runs on a copy under /tmp. The reference's pyproject.toml lists
# `packages = ["llm_quest"]`, which installs the top-level package WITHOUT its sub-packages (6 of 99 modules);
# the copy's packaging table is switched to package discovery (`include = ["llm_quest*"]`) — no source file is
# touched, and the script verifies afterwards that every installed .py is byte-identical to the reference's.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
TMP="$(mktemp -d /tmp/llmq_ref.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'EOF'
import sys
p = sys.argv[1]
s = open(p).read()
old = '[tool.setuptools]\npackages = ["llm_quest"]\npy-modules = ["config"]'
new = '[tool.setuptools]\npy-modules = ["config"]\n\n[tool.setuptools.packages.find]\ninclude = ["llm_quest*"]\nnamespaces = true'
assert old in s, "unexpected packaging table in the reference's pyproject.toml"
open(p, "w").write(s.replace(old, new))
EOF
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP" >/dev/null
find "$HERE/_ref" -name __pycache__ -prune -exec rm -rf {} \;
# unmodified? every installed module must equal the reference's
( cd "$SRC" && find llm_quest config.py -name '*.py' | sort ) | while read -r f; do
  cmp -s "$SRC/$f" "$HERE/_ref/$f" || { echo "MISMATCH $f" >&2; exit 1; }
done
echo "installed $(find "$HERE/_ref" -name '*.py' | wc -l) reference modules into $HERE/_ref (byte-identical to $SRC)"
rm -rf "$TMP"
