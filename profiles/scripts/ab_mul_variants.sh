for v in "" _s1 _s3 _s3b2 _s7; do
  export BPK_LIB=/root/repo/baby-plonk-rust_b200/libbpk$v.so
  python bench.py --logn 22 --steps 3 --warmup 2 --no-cpu > gpurun_out/ab$v.json 2> gpurun_out/ab$v.err || tail -3 gpurun_out/ab$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab$v.json"))
print("variant '$v': msm22=%.2f acc=%.2f reduce=%.2f fin=%.2f | msm20=%.2f ntt22=%.3f ntt24=%.3f"%(d["value"], d["stages_ms"]["msm.accumulate"], d["stages_ms"]["msm.reduce"], d["stages_ms"]["msm.finalize"], d["also"]["msm_ms_2^20"]["min"], d["also"]["ntt_ms_2^22"]["min"], d["also"]["ntt_ms_2^24"]["min"]))
PY
done
