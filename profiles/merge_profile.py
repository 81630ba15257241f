"""Small driver for ncu captures of the merge kernel (K3/K4):
    ncu --set full --clock-control none --import-source on -k regex:k_merge_tiles -s 1 -c 2 \
        -o gpurun_out/merge_r1 python profiles/merge_profile.py
Merges two consolidated OrdIndexedZSet<u64,u64,i64> batches of `rows` rows each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    n_val_lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    print(bench.merge_sweep(0, rows=rows, n_val_lanes=n_val_lanes))
