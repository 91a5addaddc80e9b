#!/usr/bin/env bash
# Stage the UNMODIFIED reference (pure Python) as one archive under the git-ignored oracle/_ref/, so that it travels
# to the GPU box with the gpurun snapshot (/root/reference does not exist there).  bench.py --impl reference, the
# cpu_baseline leg and the library_bar leg import it from the archive (zipimport); nothing in the product does.
# Usage: tools/stage_reference.sh [reference_dir]      (default /root/reference)
set -euo pipefail
REF="${1:-${VQA_REFERENCE:-/root/reference}}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
if [ ! -d "$REF/models" ]; then
  echo "stage_reference: $REF not present, nothing staged" >&2
  exit 0
fi
mkdir -p "$HERE/oracle/_ref"
python - "$REF" "$HERE/oracle/_ref/reference.zip" <<'PY'
import os, sys, zipfile
ref, out = sys.argv[1], sys.argv[2]
tmp = out + ".tmp"
with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
    for pkg in ("models", "api", "data", "utils", "training"):
        for root, _, files in os.walk(os.path.join(ref, pkg)):
            for f in sorted(files):
                if f.endswith(".py"):
                    full = os.path.join(root, f)
                    z.write(full, os.path.relpath(full, ref))
os.replace(tmp, out)
print("staged", out, os.path.getsize(out), "bytes")
PY
