"""BASELINE config 4: prober-gated batch.  Times the fused prober (tcgen05 bf16x3) against the only
existing GPU implementation of this step, torch-eager fp32 ImprovedProbe x6 + softmax-sum gate
(/root/reference/exp_rag.py:381-415 batched), and the gated BM25 top-10 that follows.

    python tools/bench_prober.py [--rows 16384] [--out gpurun_out/prober_bench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probing_rag_b200 import synth as po  # noqa: E402  (synthetic probers + hidden states)
from probing_rag_b200.prober import ImprovedProbe, ProberGate  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16384)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "prober_bench.json"))
    ap.add_argument("--with-bm25", action="store_true", help="also time the whole config-4 step on the full-size index")
    ap.add_argument("--n-docs", type=int, default=21015324)
    args = ap.parse_args()
    dev = torch.device("cuda")
    sds = [po.make_prober_state(l) for l in po.PROBE_LAYERS]
    gate = ProberGate(sds, device=dev)
    eager = []
    for sd in sds:
        m = ImprovedProbe(2048, 2)
        m.load_state_dict(sd)
        eager.append(m.to(dev).eval())
    x = po.make_hidden_states(args.rows, seed=4).to(dev)

    @torch.no_grad()
    def run_eager():
        acc = torch.zeros(args.rows, 2, device=dev)
        for i, m in enumerate(eager):
            acc += torch.softmax(m.forward_eager(x[:, i]), dim=1)
        return ~(acc[:, 0] + 0.0 < acc[:, 1])

    ms_fused = timeit(lambda: gate(x, sync=False))
    ms_eager = timeit(run_eager)
    out = gate(x, want_logits=True)
    ref = torch.stack([m.forward_eager(x[:, i]) for i, m in enumerate(eager)], 1)
    dp = (torch.softmax(out.logits, -1) - torch.softmax(ref, -1)).abs().max().item()
    flops = gate.flops(args.rows)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    rec = {"rows": args.rows, "fused_ms": ms_fused, "torch_eager_fp32_ms": ms_eager, "speedup": ms_eager / ms_fused,
           "queries_per_s": args.rows / ms_fused * 1e3, "algorithmic_tflops": flops / ms_fused / 1e9,
           "issued_tflops_bf16x3": 3 * flops / ms_fused / 1e9, "peak_tflops_sustained": peak,
           "frac_algorithmic": flops / ms_fused / 1e9 / peak, "frac_issued": 3 * flops / ms_fused / 1e9 / peak,
           "max_abs_dprob_vs_torch_fp32": dp, "retrieve_rate": float(out.retrieve.float().mean().item())}
    if args.with_bm25:
        # BASELINE config 4 end to end on the device: hidden states -> prober -> gate -> compaction -> BM25 top-10 of
        # the queries that retrieve, over the 21M-passage synthetic index
        import bench
        from probing_rag_b200 import BM25Retriever
        from probing_rag_b200.prober import gate_and_retrieve
        gi, qi, qt = bench.build_workload(args.n_docs, 1 << 22, args.rows, dev)
        retr = BM25Retriever.from_defaults(index=gi, similarity_top_k=10)
        d_qi, d_qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
        ms_cfg4 = timeit(lambda: gate_and_retrieve(gate, retr, x, d_qi, d_qt), reps=3, warm=2)
        n_ret = int(gate(x).retrieve.sum().item())
        rec.update({"config4_ms": ms_cfg4, "config4_queries_per_s": args.rows / ms_cfg4 * 1e3, "config4_retrieving_queries": n_ret,
                    "config4_n_docs": args.n_docs, "prober_share_of_config4": ms_fused / ms_cfg4})
    print(json.dumps(rec))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rec, f)


if __name__ == "__main__":
    main()
