"""Per-launch event counters of the flat-step kernel (instrumented -DPR_STATS variant build only):

    nvcc ... -DPR_STATS -o build_variants/lib_stats.so ; PR_LIB_PATH=build_variants/lib_stats.so python tools/skip_stats.py

Prints, for modes 6 and 7, the steps / END steps / rescored candidates / inserts per launch.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from probing_rag_b200 import _lib, synth  # noqa: E402

NAMES = ["items", "gen", "wide", "narrow", "special", "ends", "rescored", "skipped_terms", "terms", "inserts", "scan_ends"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-docs", type=int, default=synth.N_DOCS_WIKI)
    ap.add_argument("--n-queries", type=int, default=8192)
    ap.add_argument("--modes", default="6,7")
    ap.add_argument("--tune", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "skip_stats.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gi, qi, qt = bench.build_workload(args.n_docs, 1 << 22, args.n_queries, dev)
    d_qi, d_qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
    L = _lib.lib()
    buf = (ctypes.c_ulonglong * (512 * 16))()
    res = {}
    extra = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune.split(",") if kv}
    for mode in [int(m) for m in args.modes.split(",")]:
        gi.set_tuning(mode=mode, **extra)
        gi.topk(d_qi, d_qt, 10)
        torch.cuda.synchronize()
        L.pr_debug_stats(buf, 512 * 16)          # clear warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gi.topk(d_qi, d_qt, 10)
        e1.record()
        torch.cuda.synchronize()
        L.pr_debug_stats(buf, 512 * 16)
        a = np.array(buf[:], dtype=np.int64).reshape(512, 16)
        n_l = int((a[:, 0] > 0).sum())
        print(f"== mode {mode}: {e0.elapsed_time(e1):.1f} ms (instrumented), {n_l} launches, per-query totals:")
        tot = a.sum(0)
        print("   " + ", ".join(f"{n}={tot[i] / args.n_queries:.1f}" for i, n in enumerate(NAMES)))
        for li in [0, 1, 2, 3, 5, 10, 20, 50, 100, 150, 200]:
            if li < n_l:
                print(f"   launch {li:3d}: " + ", ".join(f"{n}={a[li, i] / args.n_queries:.2f}" for i, n in enumerate(NAMES)))
        res[mode] = a[:n_l, :len(NAMES)].tolist()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"names": NAMES, "n_queries": args.n_queries, "per_launch": res}, open(args.out, "w"))


if __name__ == "__main__":
    main()
