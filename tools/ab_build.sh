#!/bin/sh
# A/B helper: build libslamfe variants with different -D flags into gpurun_out-independent names.
# usage: tools/ab_build.sh name "-DTRK_MINB=5"
set -e
cd "$(dirname "$0")/../slam-robot_b200/csrc"
./build.sh $2 >/dev/null 2>&1
cp libslamfe.so libslamfe_$1.so
