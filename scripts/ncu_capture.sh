#!/bin/bash
# usage: ncu_capture.sh TAG KERNEL_REGEX SKIP -- command...   (run on the GPU box; leaves gpurun_out/TAG_{raw,sass}.csv)
TAG=$1; KRE=$2; SKIP=$3; shift 4
ncu --clock-control none --set full --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o /tmp/$TAG "$@" > /tmp/$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_src.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_src.csv gpurun_out/${TAG}_sass.csv
rm -f /tmp/$TAG.ncu-rep
