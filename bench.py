"""bench.py -- BM25 top-10 queries/sec over a 21M-passage corpus on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: all n_queries synthetic queries scored
against the whole (doc-sharded for N>1) synthetic corpus and reduced to top-k, including for
N>1 the NCCL all-gather of the per-shard lists and the merge.  Prints ONE JSON line.

  value        queries/s, inputs resident in HBM, CUDA events, max over ranks
  e2e          same through the host-buffer API (pinned H2D of the query batch, D2H of the lists)
  roofline     scoring kernel: algorithmic bytes / its summed device time
  cpu_baseline the reference's CPU algorithm (oracle port, NumPy like bm25s) on a bounded sample
  --impl reference   times only that CPU path (all host threads) and prints the same line shape
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from probing_rag_b200 import synth  # noqa: E402

METRIC = "bm25_top10_queries_per_sec"
UNIT = "queries/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- workload
from probing_rag_b200.sharding import gather_lists, global_stats, shard_range  # noqa: E402


def build_workload(n_docs: int, vocab: int, n_queries: int, device, rank: int = 0, world: int = 1,
                   query_kind: str = "round0"):
    """Synthetic Zipf-Mandelbrot corpus shard [lo, hi) built into a BM25Index on `device` with
    GLOBAL N / avgdl / df, and the seeded query batch (host CSR)."""
    import torch.distributed as dist
    from probing_rag_b200 import BM25Index
    lo, hi = shard_range(n_docs, rank, world)
    cdf = torch.from_numpy(synth.zipf_mandelbrot_cdf(vocab)).to(device)
    toks, lens = [], []
    t0 = time.time()
    for blk in range(lo // synth.DOC_BLOCK, (max(hi, 1) - 1) // synth.DOC_BLOCK + 1):
        t, l = synth.corpus_block_torch(blk, n_docs, cdf)
        b_lo = blk * synth.DOC_BLOCK
        s, e = max(lo - b_lo, 0), min(hi - b_lo, l.numel())
        if e <= s:
            continue
        if s > 0 or e < l.numel():
            off = torch.zeros(l.numel() + 1, dtype=torch.int64, device=device)
            torch.cumsum(l.long(), 0, out=off[1:])
            t = t[int(off[s]):int(off[e])]
            l = l[s:e]
        toks.append(t)
        lens.append(l)
    tokens = torch.cat(toks) if toks else torch.zeros(0, dtype=torch.int32, device=device)
    doc_lens = torch.cat(lens) if lens else torch.zeros(0, dtype=torch.int32, device=device)
    del toks, lens
    n_tok_local = tokens.numel()
    # global statistics
    from probing_rag_b200.index import bm25_weights, count_postings, idf_lucene_table
    term, doc, tf, df_local = count_postings(tokens, doc_lens, vocab)
    del tokens
    df, avgdl = global_stats(df_local, n_tok_local, n_docs)
    df_host = df.cpu().numpy()
    idf = torch.from_numpy(idf_lucene_table(df_host, n_docs)).to(device)
    w = bm25_weights(term, doc, tf, doc_lens, idf, avgdl)
    del tf, term
    indptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    torch.cumsum(df_local, 0, out=indptr[1:])
    gi = BM25Index(indptr, doc, w, hi - lo, n_docs, lo, meta={"avgdl": avgdl})
    torch.cuda.synchronize(device)
    q_indptr, q_terms = synth.queries_np(n_queries, vocab, df_host, kind=query_kind)
    log(f"[bench] rank {rank}: shard docs [{lo},{hi}) nnz={gi.nnz:,} avgdl={avgdl:.3f} "
        f"queries={n_queries} ({len(q_terms) / max(n_queries, 1):.2f} terms/query) built in {time.time() - t0:.1f}s")
    return gi, q_indptr, q_terms


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_qps(host_index: dict, q_indptr, q_terms, k: int, budget_s: float, max_queries: int,
                      n_threads: int):
    """The reference's CPU algorithm (bm25s NumPy path, SURVEY App. A.5-A.6: dense f32
    accumulator, np.add.at per query token, argpartition+argsort) via the oracle port, with
    bm25s's own thread-pool-over-queries mechanism.  Returns (queries/s, sample size)."""
    from oracle import bm25_oracle as bo
    n_probe = min(max(n_threads, 4), len(q_indptr) - 1)
    t0 = time.perf_counter()
    bo.retrieve_batch(host_index, q_indptr[:n_probe + 1], q_terms, k, n_threads=n_threads, canonical=False)
    dt = time.perf_counter() - t0
    n = int(min(max_queries, len(q_indptr) - 1, max(n_probe, budget_s / max(dt / n_probe, 1e-9))))
    t0 = time.perf_counter()
    bo.retrieve_batch(host_index, q_indptr[:n + 1], q_terms, k, n_threads=n_threads, canonical=False)
    dt = time.perf_counter() - t0
    return n / dt, n


def host_copy(gi) -> dict:
    return {"data": gi.weights.cpu().numpy(), "indices": gi.doc_ids.cpu().numpy(),
            "indptr": gi.indptr.cpu().numpy(), "num_docs": gi.n_docs, "doc_id_base": gi.doc_id_base}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-docs", type=int, default=synth.N_DOCS_WIKI)
    ap.add_argument("--vocab", type=int, default=1 << 22)
    ap.add_argument("--n-queries", type=int, default=65536)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tune", default="", help="comma list key=value for pr_bm25_tuning_t")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return                                        # rank 0 alone runs the CPU arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1 and args.impl == "ours":
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world if args.impl == "ours" else args.gpus
    workload = (f"BM25 top-{args.k} over {args.n_docs:,}-passage DPR-Wikipedia-shaped synthetic corpus, "
                f"{args.n_queries:,} queries" + (f", doc-sharded over {world} GPUs" if world > 1 else ", 1xB200"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        gi, qi, qt = build_workload(args.n_docs, args.vocab, args.n_queries, device)
        host = host_copy(gi)
        del gi
        torch.cuda.empty_cache()
        per_step_budget = max(5.0, min(40.0, 150.0 / max(args.steps + args.warmup, 1)))
        vals, n = [], 0
        for i in range(args.warmup + args.steps):
            v, n = cpu_reference_qps(host, qi, qt, args.k, per_step_budget, 512, cores)
            if i >= args.warmup:
                vals.append(v)
        v = float(np.mean(vals))
        sample = f"{n} of {args.n_queries} queries per step (seeded subsample, whole corpus)"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "k": args.k, "n_docs": args.n_docs, "n_queries": args.n_queries},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    from probing_rag_b200 import merge_topk
    gi, qi, qt = build_workload(args.n_docs, args.vocab, args.n_queries, device, rank, world)
    if args.tune:
        gi.set_tuning(**{kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune.split(",")})
    nq, k = args.n_queries, args.k
    d_qi = torch.from_numpy(qi).to(device)
    d_qt = torch.from_numpy(qt).to(device)
    out = (torch.empty((nq, k), dtype=torch.float32, device=device),
           torch.empty((nq, k), dtype=torch.int32, device=device))
    gath_s = torch.empty((world, nq, k), dtype=torch.float32, device=device) if world > 1 else None
    gath_d = torch.empty((world, nq, k), dtype=torch.int32, device=device) if world > 1 else None

    def step_device():
        gi.topk(d_qi, d_qt, k, out=out, check_status=False)
        if world > 1:
            gather_lists(out[0], out[1], out=(gath_s, gath_d))
            return merge_topk(gath_s, gath_d)
        return out

    def step_host():
        s, d, h2d, d2h = gi.topk_host(qi, qt, k)
        if world > 1:
            ds, dd = torch.from_numpy(s).to(device), torch.from_numpy(d).to(device)
            gather_lists(ds, dd, out=(gath_s, gath_d))
            ms, md = merge_topk(gath_s, gath_d)
            s, d = ms.cpu().numpy(), md.cpu().numpy()
        return s, d, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = None
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), r

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = gi.last_launches + (1 if world > 1 else 0)

    # scoring-kernel device time for the roofline (events around its launches)
    gi.set_profiling(True)
    barrier()
    step_device()
    torch.cuda.synchronize(device)
    score_ms, score_launches = gi.profile()
    gi.set_profiling(False)
    alg_bytes = gi.algorithmic_bytes(qi, qt, k)

    # end to end through the host-buffer API
    for _ in range(2):
        step_host()
    ms_e2e, last = timed(step_host, args.steps)
    _, _, h2d, d2h = last

    # size-independent sanity on the measured output (not a parity claim; tests/ hold those)
    res = step_device()
    torch.cuda.synchronize(device)
    s_chk = res[0]
    assert bool((s_chk[:, :-1] >= s_chk[:, 1:]).all()), "ranked lists not score-descending"

    if world > 1:
        t = torch.tensor([alg_bytes, score_ms * 1e3], dtype=torch.float64, device=device)
        mx = t.clone()
        dist.all_reduce(t)                       # total algorithmic bytes over shards
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        alg_total = float(t[0].item())
    else:
        alg_total = float(alg_bytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tuning = gi.get_tuning()
    kernel = {1: "bm25_score_kernel", 2: "bm25_score_kernel", 3: "bm25_warp_kernel", 4: "bm25_warp_kernel",
              8: "bm25_lean_kernel" if gi.aux_info().get("lean_ok") else "bm25_flat_kernel"}.get(
        tuning["mode"], "bm25_flat_kernel")
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/), if any
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("kernel") == kernel:
            traffic = int(tj["dram_bytes_per_launch"])
            traffic_src = tj.get("source")
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    achieved = alg_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
    value = nq * args.steps / (ms_dev * 1e-3)
    e2e = nq * args.steps / (ms_e2e * 1e-3)

    cpu_base = None
    if not args.no_cpu_baseline and world == 1:
        host = host_copy(gi)
        v, n = cpu_reference_qps(host, qi, qt, k, args.cpu_budget_s, 512, cores)
        cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{n} of {nq} queries (seeded subsample, whole corpus), NumPy oracle port, "
                              f"thread pool over queries like bm25s n_threads={cores}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "k": k, "n_docs": args.n_docs, "n_queries": nq, "vocab": args.vocab,
                   "nnz_shard0": gi.nnz, "l2": "inputs (index shard of %.1f GB) larger than L2" % (gi.nnz * 8 / 1e9),
                   "tuning": tuning, "parallelism": f"doc-shard x{world}" if world > 1 else "single"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_8000_nominal": achieved / 8000.0, "traffic": traffic,
                     "traffic_unit": "DRAM bytes (read + write) per launch, ncu", "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel": kernel,
                     "algorithmic_bytes_per_launch": int(alg_bytes / max(score_launches, 1)),
                     "algorithmic_bytes_per_step_rank0": int(alg_bytes),
                     "algorithmic_bytes_per_step_all_ranks": int(alg_total),
                     "kernel_ms_per_step": score_ms, "kernel_launches_per_step": score_launches,
                     "kernel_share_of_step": score_ms / (ms_dev / args.steps)},
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
