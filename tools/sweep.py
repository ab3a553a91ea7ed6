"""Tuning sweep of the BM25 scoring kernel on one GPU: builds the synthetic workload once and
times pr_bm25_topk for a list of pr_bm25_tuning_t settings (CUDA events, resident inputs).

    python tools/sweep.py [--n-docs N] [--n-queries B] [--k K] [--out gpurun_out/sweep.jsonl] [--grid small|full]
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
    ap.add_argument("--vocab", type=int, default=1 << 22)
    ap.add_argument("--n-queries", type=int, default=65536)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--grid", default="plan")
    ap.add_argument("--configs", default="", help="semicolon list of 'k=v,k=v' overriding --grid")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gi, qi, qt = bench.build_workload(args.n_docs, args.vocab, args.n_queries, dev)
    d_qi, d_qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
    alg = gi.algorithmic_bytes(qi, qt, args.k)
    print(json.dumps({"aux": gi.aux_info()}), flush=True)
    if args.configs:
        cfgs = [{kv.split("=")[0]: int(kv.split("=")[1]) for kv in c.split(",")} for c in args.configs.split(";")]
    else:
        cfgs = [dict(subs_per_item=g, warps_per_cta=nw, docs_per_launch=dpl)
                for g, nw, dpl in itertools.product([24, 12, 48], [8, 12], [393216, 196608, 786432, 1572864])]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    ref = None
    with open(args.out, "a") as f:
        for cfg in cfgs:
            try:
                gi.set_tuning(**cfg)
                s, d = gi.topk(d_qi, d_qt, args.k)            # warm-up + correctness cross-check
                torch.cuda.synchronize()
                if ref is None:
                    ref = (s.clone(), d.clone())
                same = bool(torch.equal(ref[0], s) and torch.equal(ref[1], d))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    gi.topk(d_qi, d_qt, args.k, check_status=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                rec = dict(cfg, ms=ms, qps=args.n_queries / ms * 1e3, alg_gbs=alg / ms / 1e6,
                           launches=gi.last_launches, same_as_first=same)
            except Exception as ex:  # keep sweeping
                rec = dict(cfg, error=str(ex))
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            f.flush()


if __name__ == "__main__":
    main()
