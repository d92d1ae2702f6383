"""Builds the C-ABI CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python graph-transformer_b200/build.py

Output: graph-transformer_b200/u2gnn_b200/libu2gnn_b200.so (git-ignored; travels to the GPU box) - the product - and
libu2gnn_b200_probe.so next to it: the same sources compiled with -DU2GNN_PROBE_BUILD plus csrc/probe/*.cu (layout
self-tests, micro-benchmarks, kernel tracing; include/u2gnn_b200_probe.h).  One translation unit per kernel family,
compiled in parallel, linked with the static CUDA runtime; no torch, no Python in the libraries.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "u2gnn_b200", "libu2gnn_b200.so")
OUT_PROBE = os.path.join(HERE, "u2gnn_b200", "libu2gnn_b200_probe.so")
OBJ = os.path.join(HERE, "build")
PROBE_SRC = os.path.join(CSRC, "probe")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha1()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = _sources()
    probe_srcs = sorted(f for f in os.listdir(PROBE_SRC) if f.endswith(".cu"))
    deps = ([os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, f))] +
            [os.path.join(PROBE_SRC, f) for f in probe_srcs] +
            [os.path.join(HERE, "..", "include", "u2gnn_b200.h"), os.path.join(HERE, "..", "include", "u2gnn_b200_probe.h")])
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(OUT) and os.path.exists(OUT_PROBE) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    os.makedirs(os.path.join(OBJ, "probe"), exist_ok=True)

    def compile_one(job):
        src, obj, extra = job
        cmd = [NVCC] + FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr)
        return obj

    jobs = [(os.path.join(CSRC, f), os.path.join(OBJ, f[:-3] + ".o"), []) for f in srcs]
    # the probe build recompiles only the units that change under U2GNN_PROBE_BUILD (the FFN kernels) and adds csrc/probe
    traced = [f for f in srcs if f.startswith("ffn_tc")]
    pjobs = [(os.path.join(CSRC, f), os.path.join(OBJ, "probe", f[:-3] + ".o"), ["-DU2GNN_PROBE_BUILD"]) for f in traced]
    pjobs += [(os.path.join(PROBE_SRC, f), os.path.join(OBJ, "probe", f[:-3] + ".o"), ["-DU2GNN_PROBE_BUILD"]) for f in probe_srcs]
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, jobs + pjobs))
    prod = objs[:len(jobs)]
    probe = [o for (f, o) in zip(srcs, prod) if f not in traced] + objs[len(jobs):]
    for out, group in ((OUT, prod), (OUT_PROBE, probe)):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + group
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
