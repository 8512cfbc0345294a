#!/bin/bash
# Round 2: per-instruction (SASS) and per-line (CUDA) ncu source pages of the level-0 batched-affine kernel at the
# headline size.  The report is converted on the box (gpurun_out/ is capped at 64 MiB) and removed.
# usage: profiles/scripts/r2_profile_source.sh <tag> [skip]
TAG=${1:-r2src}
SKIP=${2:-8}
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:msm_affine_level -s $SKIP -c 1 -o /tmp/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2> gpurun_out/${TAG}_sass.err
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_cuda.csv 2> gpurun_out/${TAG}_cuda.err
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2> gpurun_out/${TAG}_raw.err
ls -la gpurun_out/${TAG}_*
