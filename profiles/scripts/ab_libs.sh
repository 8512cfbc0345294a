#!/bin/bash
# A/B of kernel build variants on the GPU box: profiles/scripts/ab_libs.sh <tag> <variant> [<variant> ...] [-- extra bench args]
# variant "default" = baby-plonk-rust_b200/libbpk.so, otherwise baby-plonk-rust_b200/libbpk_<variant>.so
TAG=$1; shift
EXTRA=""
VARS=()
while [ $# -gt 0 ]; do
  if [ "$1" == "--" ]; then shift; EXTRA="$@"; break; fi
  VARS+=("$1"); shift
done
mkdir -p gpurun_out
for v in "${VARS[@]}"; do
  lib=baby-plonk-rust_b200/libbpk.so
  [ "$v" != "default" ] && lib=$PWD/baby-plonk-rust_b200/libbpk_$v.so
  BPK_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 2 --no-extras --no-cpu --no-verify $EXTRA > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$v.json"))
    print("$v", "value %.2f e2e %.2f" % (d["value"], d["e2e"]["value"]), d["stages_ms"], "exec_frac %.3f" % (d["roofline"]["executed_frac"] or 0))
except Exception as e:
    print("$v FAILED", e, open("gpurun_out/${TAG}_$v.err").read()[-500:])
PY
done
