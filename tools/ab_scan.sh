#!/bin/bash
# development aid: same-box A/B of two builds of the library over the scan + strip sweep shapes
# usage: tools/ab_scan.sh "64 256 1024 DENSE" libA.so libB.so
sizes=$1; shift
for lib in "$@"; do
  echo "== $lib"
  HEVCB_LIB=$PWD/$lib GIB=${GIB:-1} bash tools/sweep_quick.sh "$sizes" ""
done
