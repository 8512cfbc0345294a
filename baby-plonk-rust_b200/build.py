"""Build libbpk.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python baby-plonk-rust_b200/build.py [--force]

The shared library is written next to this file (baby-plonk-rust_b200/libbpk.so) so that it travels
to the GPU box with the repository snapshot.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libbpk.so")
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "msm_tree.cu", "srs.cu", "polyops.cu"]
HEADERS = ["ff.cuh", "ec.cuh", "internal.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_dep():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "bpk.h"))
    deps.append(os.path.abspath(__file__))
    return max(os.path.getmtime(d) for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    log = os.path.join(OBJ, src.replace(".cu", ".ptxas.log"))
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
    return obj


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        for src in SOURCES:
            print(open(os.path.join(OBJ, src.replace(".cu", ".ptxas.log"))).read())
    return LIB


def build_variant(tag, defines):
    """Experimental build with extra -D flags -> libbpk_<tag>.so (loaded with BPK_LIB=<path>).
    Used for A/B measurements of kernel formulations on the GPU box; not the shipped library."""
    vobj = os.path.join(HERE, "_obj_" + tag)
    os.makedirs(vobj, exist_ok=True)
    out = os.path.join(HERE, "libbpk_%s.so" % tag)

    def comp(src):
        obj = os.path.join(vobj, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(vobj, src.replace(".cu", ".ptxas.log")), "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(comp, SOURCES))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
