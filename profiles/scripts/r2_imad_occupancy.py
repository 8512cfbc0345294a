#!/usr/bin/env python3
"""Duty cycle of the wide-IMAD pipe vs resident warps per SM, for the register-only probes and for the real Fp multiplier
(bpk_imad_peak modes: 1 = two 12-limb carry chains, 2 = 14 independent fused accumulates, 3 = Fp product through the
out-of-line body the MSM kernels call, 4 = the same product inlined, 5 = two independent inlined products).
Run under gpurun: python profiles/scripts/r2_imad_occupancy.py > gpurun_out/r2_imad_occupancy.json"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("baby-plonk-rust_b200")
ctx = pkg.get_context()
rows = []
for warps in (8, 12, 16, 20, 24, 32, 48, 64):
    ctx.set_option("imad.warps_per_sm", warps)
    row = {"warps_per_sm": warps}
    for mode in (1, 2, 3, 4, 5):
        best = 0.0
        for _ in range(3):
            rate, _ = ctx.imad_peak(mode)
            best = max(best, rate)
        row["mode%d_wide_imad_per_s" % mode] = best
    rows.append(row)
    print(json.dumps(row), flush=True)
ctx.set_option("imad.warps_per_sm", 64)
