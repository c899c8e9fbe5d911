#!/usr/bin/env bash
# Test infrastructure, never shipped: copies the UNMODIFIED pure-Python reference (Wall-Go/WallGo) and the four
# example drivers BASELINE.json's configs name into the git-ignored oracle/_ref/, so that the side-by-side parity
# tests, the end-to-end solveWall gate and `bench.py --impl reference` can run the reference itself on the GPU
# box (oracle/_ref/ is git-ignored but NOT gpurun-ignored: it travels with the snapshot like the built .so).
# Nothing under wallgo_b200/ imports it.  Collision .hdf5 files are Git-LFS pointers and are not copied.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/src/WallGo" ]; then
  echo "fetch_ref: $SRC/src/WallGo not found (nothing to do on this machine)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/src" "$DST/Models"
cp -r "$SRC/src/WallGo" "$DST/src/WallGo"
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
cp "$SRC/Models/wallGoExampleBase.py" "$DST/Models/"
for m in Yukawa SingletStandardModel_Z2 StandardModel InertDoubletModel; do
  mkdir -p "$DST/Models/$m"
  find "$SRC/Models/$m" -maxdepth 1 -type f \( -name '*.py' -o -name '*.ini' -o -name '*.txt' \) -exec cp {} "$DST/Models/$m/" \;
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "v1.1.2 (no git metadata)" ) > "$DST/VERSION"
echo "fetched reference into $DST ($(du -sh "$DST" | cut -f1))"
