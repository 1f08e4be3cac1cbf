#!/bin/sh
# build a named library variant: tools/ab_build.sh <name> [extra nvcc flags]; fails loudly when the build does
set -e
cd "$(dirname "$0")/../slam-robot_b200/csrc"
name=$1; shift
rm -f libslamfe.so
./build.sh "$@" > /tmp/ab_build_$name.log 2>&1 || { grep -E "error" /tmp/ab_build_$name.log; echo "BUILD FAILED: $name"; exit 1; }
grep -A2 "track_fb_kernel" /tmp/ab_build_$name.log | grep -E "registers|spill" | tr '\n' ' '; echo
cp libslamfe.so libslamfe_$name.so
echo "built $name"
