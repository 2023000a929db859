#!/bin/bash
# One `ncu --set full` capture of a workload of scripts/ncu_families.py on the GPU box, reduced ON THE BOX to the raw-page
# CSV (a 25-kernel report is ~140 MB, over gpurun's 64 MiB return limit):
#   scripts/ncu_capture.sh <tag> <workload> [ncu options, e.g. -k regex:lstm_fused -c 3]
# writes gpurun_out/<tag>.csv.gz (every metric of every captured launch) and gpurun_out/<tag>.log
set -u
tag=$1; shift
workload=$1; shift
mkdir -p gpurun_out
python scripts/ncu_families.py "$workload" > "gpurun_out/${tag}_plain.log" 2>&1 || { echo "plain run failed"; tail -5 "gpurun_out/${tag}_plain.log"; exit 1; }
ncu --set full --clock-control none --profile-from-start off "$@" -f -o "/tmp/${tag}" python scripts/ncu_families.py "$workload" > "gpurun_out/${tag}.log" 2>&1
echo "ncu rc=$?" >> "gpurun_out/${tag}.log"
if [ -f "/tmp/${tag}.ncu-rep" ]; then
  ncu -i "/tmp/${tag}.ncu-rep" --page raw --csv 2>/dev/null | gzip -9 > "gpurun_out/${tag}.csv.gz"
  ls -la "/tmp/${tag}.ncu-rep" "gpurun_out/${tag}.csv.gz" >> "gpurun_out/${tag}.log"
fi
tail -2 "gpurun_out/${tag}.log"
