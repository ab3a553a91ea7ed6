"""Small-batch latency of pr_bm25_topk on the full-size index (the reference calls retrieve() one query at a time).

    python tools/latency.py [--n-docs N] [--batches 1,8,64,512,4096] [--k 10]
"""
import argparse
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
    ap.add_argument("--batches", default="1,8,64,512,4096")
    ap.add_argument("--k", default="10", help="comma list of depths")
    ap.add_argument("--kind", default="round0", help="round0 (question-sized queries) or later (LM-transcript-sized, 64..1024 terms)")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--graph", action="store_true", help="also time the calls replayed from a CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bs = [int(x) for x in args.batches.split(",")]
    gi, qi, qt = bench.build_workload(args.n_docs, args.vocab, max(bs), dev, query_kind=args.kind)
    for b, k in [(b, int(k)) for b in bs for k in args.k.split(",")]:
        # a "batch" of b queries; for b = 1 the mean over up to 64 DIFFERENT single queries, one call each (one query's
        # time is that of its terms' posting lists: a single sample says little)
        n_calls = min(64, len(qi) - 1) if b == 1 else 1
        calls = []
        for c in range(n_calls):
            lo, hi = (c, c + 1) if b == 1 else (0, b)
            calls.append((torch.from_numpy(qi[lo:hi + 1] - qi[lo]).to(dev), torch.from_numpy(qt[qi[lo]:qi[hi]]).to(dev)))
        for _ in range(3 if b > 1 else 1):
            for d_qi, d_qt in calls:
                gi.topk(d_qi, d_qt, k)
        torch.cuda.synchronize()
        reps = args.reps if b <= 4096 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for d_qi, d_qt in calls:
                gi.topk(d_qi, d_qt, k, check_status=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * n_calls)
        n_terms = float(qi[n_calls] if b == 1 else qi[b]) / (n_calls if b == 1 else b)
        row = {"batch": b, "k": k, "kind": args.kind, "terms_per_query": n_terms, "ms_per_call": ms, "qps": b / ms * 1e3,
               "launches": gi.last_launches, "distinct_calls_averaged": n_calls}
        if args.graph and b <= 512:
            # the same calls captured once in a CUDA graph and replayed: the device time of a call without the host's
            # per-call work (Python argument checks + ctypes + three kernel launches, which exceed the device time of
            # a single query)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    outs = [gi.topk(d_qi, d_qt, k, check_status=False) for d_qi, d_qt in calls]
                graph.replay()
                side.synchronize()
                want = [gi.topk(d_qi, d_qt, k) for d_qi, d_qt in calls]
                side.synchronize()
                assert all(torch.equal(o[0], w[0]) and torch.equal(o[1], w[1]) for o, w in zip(outs, want)), "graph replay differs"
                e0.record(side)
                for _ in range(reps):
                    graph.replay()
                e1.record(side)
                side.synchronize()
            row["ms_per_call_graph_replay"] = e0.elapsed_time(e1) / (reps * n_calls)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
