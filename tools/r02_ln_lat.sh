#!/bin/bash
cd "$(dirname "$0")/.."
for f in 0 1; do for s in 2 4 8; do [ $f = 0 ] && [ $s != 4 ] && continue
echo "== FUSE_LN=$f slice=$s"; DINO_B200_FUSE_LN=$f DINO_B200_LN_SLICE=$s python tools/latency.py vitl14 vits14 2>&1 | grep "batch 1"; done; done
