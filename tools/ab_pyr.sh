#!/bin/sh
# pyramid A/B on the GPU box: tools/ab_pyr.sh variant... (runs tools/pyr_bench.py at 128/256/512 frames per variant)
cd "$(dirname "$0")/.."
for v in "$@"; do
  cp slam-robot_b200/csrc/libslamfe_$v.so slam-robot_b200/csrc/libslamfe.so
  touch slam-robot_b200/csrc/libslamfe.so
  echo "== $v"
  for b in 128 256 512; do python tools/pyr_bench.py $b; done
done
