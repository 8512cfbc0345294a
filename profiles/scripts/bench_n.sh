#!/bin/bash
# bench.py on N GPUs of one box the way the driver launches it; prints the headline fields.  usage: bench_n.sh N tag [bench args]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
if [ "$N" == "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
fi
grep "^{" gpurun_out/${TAG}_bench$N.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); a=d.get('also',{}); p=a.get('prove_s_2^20_gates',{})
print(d['n_gpus'], 'msm %.2f e2e %.2f' % (d['value'], d['e2e']['value']), d['verified'], d['stages_ms'], d['roofline']['executed_plan'])
print('prove', p.get('prove_s'), p.get('prove_cached_s'), p.get('proof_verified'))
"
