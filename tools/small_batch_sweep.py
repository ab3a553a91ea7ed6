"""Launch-plan sweep for small batches on the full-size index: items_per_warp x subs_per_item x warps_per_cta.

    python tools/small_batch_sweep.py [--batches 1,8,64,512] [--out gpurun_out/small_sweep.jsonl]
"""
import argparse
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from probing_rag_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-docs", type=int, default=synth.N_DOCS_WIKI)
    ap.add_argument("--batches", default="1,8,64,512")
    ap.add_argument("--ipw", default="1,2,4,8,16")
    ap.add_argument("--nw", default="8")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "small_sweep.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bs = [int(x) for x in args.batches.split(",")]
    gi, qi, qt = bench.build_workload(args.n_docs, 1 << 22, max(bs), dev)
    with open(args.out, "a") as f:
        for b in bs:
            d_qi = torch.from_numpy(qi[:b + 1]).to(dev)
            d_qt = torch.from_numpy(qt[:qi[b]]).to(dev)
            alg = gi.algorithmic_bytes(qi[:b + 1], qt[:qi[b]], args.k)
            ref = None
            for ipw, nw in itertools.product([int(x) for x in args.ipw.split(",")], [int(x) for x in args.nw.split(",")]):
                gi.set_tuning(items_per_warp=ipw, warps_per_cta=nw)
                for _ in range(3):
                    s, d = gi.topk(d_qi, d_qt, args.k)
                if ref is None:
                    ref = (s.clone(), d.clone())
                same = bool(torch.equal(s, ref[0]) and torch.equal(d, ref[1]))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    gi.topk(d_qi, d_qt, args.k, check_status=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                rec = {"batch": b, "items_per_warp": ipw, "warps_per_cta": nw, "ms": ms, "qps": b / ms * 1e3,
                       "alg_gbs": alg / ms / 1e6, "launches": gi.last_launches, "same": same}
                print(json.dumps(rec), flush=True)
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
