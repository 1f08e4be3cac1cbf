#!/bin/sh
# band sweep of the pyramid build on the GPU box: tools/ab_bands.sh variant "batches" "bands"
cd "$(dirname "$0")/.."
cp slam-robot_b200/csrc/libslamfe_$1.so slam-robot_b200/csrc/libslamfe.so
touch slam-robot_b200/csrc/libslamfe.so
for b in $2; do for n in $3; do echo -n "$1 bands=$n "; SFE_PYR_BANDS=$n python tools/pyr_bench.py $b | cut -c18-; done; done
