#!/bin/bash
# The round-end sequence on one GPU: smoke, the whole GPU suite, the reference arm, the default bench.  usage: validate_all.sh <tag>
TAG=${1:-val}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 300 gpurun_out/${TAG}_ref.json; echo
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
grep "^{" gpurun_out/${TAG}_bench.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); a=d['also']
print(d['value'], d['e2e']['value'], d['verified'], d['stages_ms'], d['gpu_launches'], d['clocks'])
print({k:(v['value'], v.get('all_ms')) for k,v in d['e2e_variants'].items()})
print(d['roofline']['executed_frac'], d['roofline']['frac'], d['roofline']['traffic'])
for k in ('msm_ms_2^16','msm_ms_2^20'): print(k, a[k]['mean'])
print(a['ntt_ms_2^22']['mean'], a['prove_s_2^20_gates']['prove_s'], a['prove_s_2^20_gates'].get('prove_cached_s'), a['prove_s_2^20_gates']['proof_verified'])
"
