"""Builds compile-time variants of the merge kernel for on-GPU sweeps (how the
round-1 tile shape, register caps, look-back width and prefetch distance were
chosen):

    python profiles/sweep_build.py "MERGE_IPT_CFG=7,9,11" "MERGE_MIN_CTAS=4,5,6"   # cartesian product
    gpurun -- 'bash profiles/sweep_run.sh 50000000 1'                              # rows per input, value lanes

Every variant is merge.cu recompiled with the given -D overrides and linked with
the library's other objects into scratch/libdbsp_<variant>.so (scratch/ is
git-ignored but travels to the GPU box); sweep_run.sh points DBSP_B200_LIB at
each of them and runs profiles/merge_profile.py.  Tunables: MERGE_THREADS_CFG,
MERGE_IPT_CFG / _MID / _WIDE, MERGE_MIN_CTAS / MERGE_CTAS_MID / _WIDE,
MERGE_NARROW_MAX, MERGE_LB_THREADS, MERGE_L2_PREFETCH (see merge.cu)."""
import itertools
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main(argv):
    axes = []
    for a in argv:
        name, vals = a.split("=")
        axes.append([(name, v) for v in vals.split(",")])
    variants = {"_".join(f"{n.replace('MERGE_', '').lower()}{v}" for n, v in combo): dict(combo) for combo in itertools.product(*axes)}
    g.build()
    out_dir = os.path.join(ROOT, "scratch")
    os.makedirs(out_dir, exist_ok=True)
    objdir = os.path.join(g.CSRC, "build")
    others = [os.path.join(objdir, s.replace(".cu", ".o")) for s in g.CU_SOURCES if s != "merge.cu"]

    def one(item):
        name, defs = item
        obj = os.path.join(out_dir, f"merge_{name}.o")
        lib = os.path.join(out_dir, f"libdbsp_{name}.so")
        flags = [f"-D{k}={v}" for k, v in defs.items()]
        subprocess.check_call([g._nvcc()] + g.NVCC_FLAGS + flags + ["-c", os.path.join(g.CSRC, "merge.cu"), "-o", obj])
        subprocess.check_call([g._nvcc(), "-shared", "-o", lib, obj] + others + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
        os.remove(obj)
        return lib

    with ThreadPoolExecutor(max_workers=8) as ex:
        for lib in ex.map(one, variants.items()):
            print("built", lib)


if __name__ == "__main__":
    main(sys.argv[1:])
