#!/bin/bash
# Round 2, end of round: launch list of the default bench + ncu --set full of the MSM tree levels, the sort kernels
# and the NTT passes (run under gpurun, 1 GPU).  usage: profiles/scripts/r2_profile_final.sh <tag>
TAG=${1:-r2z}
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify"
NTT="python profiles/scripts/ntt_only.py 22"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# the timed MSM is the second one: skip the levels of the warm-up MSM
ncu --set full --clock-control none -k regex:msm_affine_level -s 8 -c 8 -o gpurun_out/${TAG}_affine $CMD > gpurun_out/${TAG}_ncu_affine.log 2>&1
# one MSM launches 1 count + 4 scatter kernels and 15 plane-tree levels: again the second MSM's
ncu --set full --clock-control none -k regex:"msm_count_kernel|msm_scatter_range_kernel" -s 5 -c 5 -o gpurun_out/${TAG}_sort $CMD > gpurun_out/${TAG}_ncu_sort.log 2>&1
ncu --set full --clock-control none -k regex:msm_plane_tree_level -s 15 -c 4 -o gpurun_out/${TAG}_tree $CMD > gpurun_out/${TAG}_ncu_tree.log 2>&1
if [ -z "$SKIP_NTT" ]; then
$NTT > gpurun_out/${TAG}_ntt_plain.log 2>&1 && ncu --set full --clock-control none -k regex:ntt_pass_r8 -s 9 -c 3 -o gpurun_out/${TAG}_ntt $NTT > gpurun_out/${TAG}_ncu_ntt.log 2>&1
fi
# the reports are large: keep their raw pages as CSV (the summaries under profiles/ are made from these), drop the rest
for r in affine sort tree ntt; do
  if [ -f gpurun_out/${TAG}_$r.ncu-rep ]; then
    ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
    rm -f gpurun_out/${TAG}_$r.ncu-rep
  fi
done
ls -la gpurun_out/${TAG}_*
