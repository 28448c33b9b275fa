#!/bin/bash
# ncu evidence of a round (run under gpurun): launch list of the cfg-2 short run + one --set full pass over the hot kernels.
# Usage: tools/profile_round.sh <tag>   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.csv
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python tools/short_run.py 60 > gpurun_out/${TAG}_short.log 2>&1 || { echo "short run failed"; tail -5 gpurun_out/${TAG}_short.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/short_run.py 60 > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"odom_loop|odom_refresh_pruned|odom_iter_cluster|sr_select|sr_curv|vox_small|vox_split|map_gn|csr_|vm_merge|rs_cluster" \
    -s 60 -c 36 -o gpurun_out/${TAG}_full -f python tools/short_run.py 30 > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | head -20
