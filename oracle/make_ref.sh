#!/usr/bin/env bash
# oracle/make_ref.sh -- stage the UNMODIFIED Python reference for the GPU box.
#
# The reference (/root/reference, pure Python) exists only in the build container; bench.py's
# `--impl reference` arm must time it on the GPU box's host cores.  This recipe copies the
# reference package and the three data files its hot path reads into the git-ignored directory
# oracle/_ref/ (listed in .gitignore, NOT in .gpurunignore: it travels with the snapshot like our
# own built .so files; it never enters the history).  Nothing is edited: `diff -r` against
# /root/reference/dpLGAR is empty, and oracle/_ref/MANIFEST records the sha256 of every file.
# Test infrastructure only: bench.py's reference arm and tests may execute it, the product never.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/dpLGAR" ]; then
  echo "make_ref.sh: $SRC/dpLGAR not found (GPU box?): keeping the prebuilt $DST" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/data"
cp -r "$SRC/dpLGAR" "$DST/dpLGAR"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
for f in vG_default_params.dat forcing_data_resampled_uniform_Phillipsburg.csv forcing_data_resampled_uniform_Bushland.csv; do
  cp "$SRC/data/$f" "$DST/data/$f"
done
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
(cd "$DST" && find . -type f ! -name MANIFEST | sort | xargs sha256sum) > "$DST/MANIFEST"
diff -r -x __pycache__ "$SRC/dpLGAR" "$DST/dpLGAR" > /dev/null && echo "oracle/_ref staged: identical to $SRC/dpLGAR ($(wc -l < "$DST/MANIFEST") files)"
