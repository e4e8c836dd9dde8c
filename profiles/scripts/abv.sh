#!/bin/bash
for f in rust-softrender_b200/csrc/libsoftrender_b200.so rust-softrender_b200/csrc/variants/*.so; do
  echo "== $f"
  SOFTRENDER_B200_LIB=$PWD/$f python profiles/scripts/earlyz.py ${1:-grid10m} 2>&1 | grep -o "reverse [A-Za-z]* precheck [0-9]\|'vertex_ms': [0-9.]*\|'micro_ms': [0-9.]*\|'raster_ms': [0-9.]*\|'total_ms': [0-9.]*\|identical [A-Za-z]*" | paste - - - - - -
done
