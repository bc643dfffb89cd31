#!/bin/bash
# quick timing of the scan + strip kernel over the sweep shapes (development aid; bench.py is the measurement)
# usage: tools/sweep_quick.sh "64 256 1024" SETTING...     (DENSE = EPB-dense 4 KiB NALs)
sizes=$1; shift
for n in $sizes; do
  if [ "$n" = DENSE ]; then echo DENSE; NAL=4096 DENSE=1 GIB=${GIB:-2} python tools/quickbench_scan.py "$@"; else echo "NAL=$n"; NAL=$n GIB=${GIB:-2} python tools/quickbench_scan.py "$@"; fi
done
