"""Time the C++ device-resident prover (tests/cpp/device_prover_main.cpp) on the chain circuit of 2^LOGN rows and
check its proof against the Python prover's: python profiles/scripts/cpp_prove_time.py LOGN"""
import importlib, os, struct, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
bpk = importlib.import_module("baby-plonk-rust_b200")
prover_mod = importlib.import_module("baby-plonk-rust_b200.prover")
synthetic = importlib.import_module("baby-plonk-rust_b200.synthetic")
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cache = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << logn
circ = synthetic.chain_circuit(n, n - 3, seed=2)
blinding = list(range(11, 22))
path = os.path.join(tempfile.gettempdir(), "bpk_instance_%d.bin" % logn)
with open(path, "wb") as f:
    f.write(struct.pack("<5Q", n, 1, n + 8, cache, 4))
    f.write(bpk.scalars_from_ints([101]).tobytes())
    for c in circ["selectors"] + circ["sigmas"] + circ["wires"]:
        f.write(np.ascontiguousarray(c).tobytes())
    f.write(bpk.scalars_from_ints(circ["public_inputs"]).tobytes())
    f.write(bpk.scalars_from_ints(blinding).tobytes())
r = subprocess.run([ge.build_cpp_host_tests("device_prover_main"), path], capture_output=True, text=True)
print("C++ prove_ms:", [l.split()[1] for l in r.stderr.splitlines() if l.startswith("prove_ms")], "rc", r.returncode)
ctx = bpk.Context(0)
setup = bpk.Setup.generate_srs(n + 8, 101, ctx)
py = prover_mod.DeviceProver(setup, n, circ["selectors"], circ["sigmas"]).prove(circ["wires"], circ["public_inputs"], blinding)
print("same proof as the Python prover:", r.stdout.split()[0] == py.to_bytes().hex())
os.remove(path)
