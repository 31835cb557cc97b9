"""In-tree build of the sm_100a shared library (plain nvcc, no JIT, no torch headers).

    python -m memflow_msci_project_b200.build          # or __graft_entry__.build()

The library is written to memflow_msci_project_b200/_lib/libmemflow_b200.so; it is git-ignored
but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libmemflow_b200.so")
SOURCES = ["abi.cu", "rambo.cu", "rambo_bwd.cu", "flow.cu", "reduce.cu", "flow_tc.cu", "rqs.cu", "flow_mma.cu", "gemm_tc.cu", "mmd.cu", "periodic.cu"]
# micro-benchmarks / building-block checks of the tcgen05 pieces: a separate TEST-ONLY library
# (tests/test_tc_probe_gpu.py, tools/probe_*.py); nothing of it is linked into the product library
PROBE_LIB = os.path.join(OUT_DIR, "libmemflow_b200_probe.so")
PROBE_SOURCES = ["abi.cu", "tc_probe.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
    "-ccbin", "/usr/bin/g++",
]


# rambo.cu: --use_fast_math only changes FLOAT arithmetic (approximate division / sqrt / transcendentals),
# i.e. the fp32 throughput mode; the fp64 parity mode is unaffected.
EXTRA_FLAGS = {"rambo.cu": ["--use_fast_math"]}
if os.environ.get("MF_TC_TRACE") == "1":  # cycle-level trace of the tensor-core flow kernel (tools/trace_tc.py)
    EXTRA_FLAGS["flow_tc.cu"] = ["-DMF_TC_TRACE"]
if os.environ.get("MF_TC_DEFS"):  # experiment switches of flow_tc.cu (comma-separated macro names), tools/gpu_iter.sh
    EXTRA_FLAGS["flow_tc.cu"] = EXTRA_FLAGS.get("flow_tc.cu", []) + ["-D" + d for d in os.environ["MF_TC_DEFS"].split(",")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "memflow_b200.h"))
    objs, jobs = [], []
    for src in SOURCES + [p for p in PROBE_SOURCES if p not in SOURCES]:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + EXTRA_FLAGS.get(os.path.basename(s), []) + ["-c", s, "-o", o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OUT_DIR, os.path.basename(s) + ".ptxas.log")
        with open(log, "w") as f:
            f.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return o

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, jobs))
    obj_of = lambda src: os.path.join(OUT_DIR, src.replace(".cu", ".o"))
    for lib, srcs in ((LIB, SOURCES), (PROBE_LIB, PROBE_SOURCES)):
        lobjs = [obj_of(x) for x in srcs]
        if jobs or force or _stale(lib, lobjs):
            cmd = [NVCC, "-shared", "-o", lib] + lobjs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                          "-ccbin", "/usr/bin/g++", "-lcudart"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
