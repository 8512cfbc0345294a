# sweep MSM plan parameters at 2^24 (precomputed SRS levels) -- prints value / accumulate / reduce ms
for args in "--fanin 4" "--fanin 8" "--fanin 16" "--fanin 32" "--pre-window 21" "--pre-window 23" "--pre-window 20" "--chunk 128" "--chunk 512" "--chunk 64"; do
  python bench.py --no-cpu --no-extras --steps 3 --warmup 2 $args 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
s = d['stages_ms']
print('$args', 'value=%.2f acc=%.2f sort=%.2f merge=%.2f reduce=%.2f fin=%.2f plan=%s' % (d['value'], s['msm.accumulate'], s['msm.sort'], s['msm.merge'], s['msm.reduce'], s['msm.finalize'], d['roofline']['executed_plan']))"
done
