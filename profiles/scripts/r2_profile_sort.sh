#!/bin/bash
# Round 2: ncu --set full of the bucket-sort kernels (count + phased scatter) of one 2^24 MSM; raw page only.
# usage: profiles/scripts/r2_profile_sort.sh <tag>
TAG=${1:-r2sort}
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
# the warm-up MSM launches 1 count + 4 scatter kernels: skip them, take the timed MSM's five
ncu --set full --clock-control none -k regex:"msm_count_kernel|msm_scatter_range_kernel" -s 5 -c 5 -o /tmp/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2> gpurun_out/${TAG}_raw.err
ls -la gpurun_out/${TAG}_*
