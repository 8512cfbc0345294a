#!/bin/bash
# Round 2: launch list + ncu --set full of the MSM kernels at the headline size (run under gpurun, 1 GPU).
# usage: profiles/scripts/r2_profile_msm.sh <tag>
TAG=${1:-r2a}
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_affine_level -s 8 -c 2 -o gpurun_out/${TAG}_affine $CMD > gpurun_out/${TAG}_ncu_affine.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"msm_scatter|msm_count" -s 2 -c 2 -o gpurun_out/${TAG}_sort $CMD > gpurun_out/${TAG}_ncu_sort.log 2>&1
ls -la gpurun_out/${TAG}_*
