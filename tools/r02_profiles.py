"""Turn the raw outputs of tools/gpu_r2_evidence.sh (gpurun_out/r2e_*) into the committed artefacts under profiles/r02
and profiles/traffic.json (read by bench.py for roofline.traffic / limiter / dram_frac).

    python tools/r02_profiles.py
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02")
G = os.path.join(ROOT, "gpurun_out")


def launch_rows(path):
    hdr, rows = None, []
    for r in csv.reader(open(path)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            rows.append(dict(zip(hdr, r)))
    return rows


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")


def launches_summary():
    rows = launch_rows(os.path.join(G, "r2e_launches_bench.csv"))
    # split into pr_bm25_topk calls at every init kernel; the second call with 100+ launches is one bench step
    calls, cur = [], []
    for r in rows:
        if "bm25_init" in r["Kernel Name"] and cur:
            calls.append(cur)
            cur = []
        cur.append(r)
    calls.append(cur)
    big = [c for c in calls if sum("bm25_lean" in r["Kernel Name"] for r in c) > 20]
    step = big[1] if len(big) > 1 else big[0]
    agg = {}
    for r in step:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    total = sum(v[1] for v in agg.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"bm25_|prober_|topk\" (python bench.py --steps 1 --warmup 1 "
             "--no-cpu-baseline), round-2 build",
             f"{len(rows)} launches of our kernels in the whole command ({len(calls)} pr_bm25_topk calls + the secondary block); below: one bench step",
             "at 21,015,324 docs x 65,536 queries; per-launch times are cold-cache and serialised under ncu", ""]
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{k:60s} launches={n:4d} total_ms={ms:9.3f} avg_us={1e3 * ms / n:9.1f} share={100 * ms / total:5.1f}%")
    sc = [float(r["Metric Value"]) / 1e6 for r in step if "bm25_lean" in r["Kernel Name"]]
    names = [short(r["Kernel Name"]) for r in step if "bm25_lean" in r["Kernel Name"]]
    lines += ["", "scoring launches of that step in order (ms; variant 0 = two tile epochs while bounds are weak, 2 = four):",
              " ".join(f"{t:.2f}{'' if 'E, 2>' in n or ', 2>' in n else '*'}" for t, n in zip(sc, names)),
              "(* = two-epoch variant)"]
    # every other kernel of ours seen in the command (prober, merges of the secondary block)
    other = {}
    for r in rows:
        k = short(r["Kernel Name"])
        a = other.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    lines += ["", "all kernels of ours in the whole command:"]
    for k, (n, ms) in sorted(other.items(), key=lambda x: -x[1][1]):
        lines.append(f"{k:60s} launches={n:5d} total_ms={ms:10.3f}")
    open(os.path.join(OUT, "launches_21M_64k_summary.txt"), "w").write("\n".join(lines) + "\n")
    import shutil
    shutil.copy(os.path.join(G, "r2e_launches_bench.csv"), os.path.join(OUT, "launches_21M_64k.csv"))
    return sum(sc), len(sc)


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def num(x):
    return float(str(x).replace(",", ""))


def traffic():
    m = raw_metrics(os.path.join(G, "r2e_prof_lean_big.ncu-rep"))[0]
    rd, wr = num(m["dram__bytes_read.sum"]), num(m["dram__bytes_write.sum"])
    # ncu prints with units chosen per column: read the unit row
    out = subprocess.run(["ncu", "-i", os.path.join(G, "r2e_prof_lean_big.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    units = dict(zip(rows[0], rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd *= scale[units["dram__bytes_read.sum"]]
    wr *= scale[units["dram__bytes_write.sum"]]
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    ms = num(m["gpu__time_duration.sum"]) * tscale[units["gpu__time_duration.sum"]]
    t = {"kernel": "bm25_lean_kernel", "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
         "captured_launch_docs": 393216, "captured_launch_ms": ms,
         "limiter": {"unit": "l1tex (shared-memory data pipe: the scatter of postings into the per-warp score tile)",
                     "l1tex_throughput_pct": num(m["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]),
                     "issue_active_pct": num(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                     "lts_throughput_pct": num(m["lts__throughput.avg.pct_of_peak_sustained_elapsed"]),
                     "dram_throughput_pct": num(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
                     "l2_hit_rate_pct": num(m["lts__t_sector_hit_rate.pct"]),
                     "shared_wavefronts": int(num(m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"])),
                     "shared_bank_conflict_wavefronts": int(num(m["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]))},
         "source": "profiles/r02/lean_large_batch_ncu_summary.txt: ncu --set full --clock-control none, one full-size launch "
                   "(393,216 documents x 65,536 queries, 46.5 GB of algorithmic bytes, four-epoch variant) in the middle of a bench.py "
                   "step at 21,015,324 docs, round 2"}
    json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    return t


def config5():
    rows = [json.loads(l) for l in open(os.path.join(G, "r2e_config5_sweep.jsonl")) if l.startswith("{")]
    import shutil
    shutil.copy(os.path.join(G, "r2e_config5_sweep.jsonl"), os.path.join(OUT, "config5_sweep.jsonl"))
    md = ["# BASELINE config 5: batch x depth x retrieval round (21,015,324 passages, 1xB200, device time per call)", "",
          "`tools/latency.py` on the full-size index; round 0 = question-sized queries (~6 terms), rounds 1-3 = the decoded",
          "transcript as the query (`exp_rag.py:428`, 64..1024 terms, ~350 on average).  ms per `pr_bm25_topk` call (queries/s);",
          "batch 1 = the mean over 64 DIFFERENT single queries, one call each (a single query's time is that of its terms'",
          "posting lists).  In brackets where measured: the same calls replayed from a CUDA graph (no per-call host work).", ""]
    for kind, title in (("round0", "round 0"), ("later", "rounds 1-3 (transcript-sized queries)")):
        sel = [r for r in rows if r["kind"] == kind]
        ks = sorted({r["k"] for r in sel})
        bs = sorted({r["batch"] for r in sel})
        md += [f"## {title}", "", "| batch | " + " | ".join(f"k={k}" for k in ks) + " |", "|---|" + "---|" * len(ks)]
        for b in bs:
            cells = []
            for k in ks:
                r = next((r for r in sel if r["batch"] == b and r["k"] == k), None)
                g = f" [{r['ms_per_call_graph_replay']:.3f}]" if r and "ms_per_call_graph_replay" in r else ""
                cells.append(f"{r['ms_per_call']:.3f} ms ({r['qps']:,.0f}/s){g}" if r else "")
            md.append(f"| {b} | " + " | ".join(cells) + " |")
        md.append("")
    open(os.path.join(OUT, "config5_sweep.md"), "w").write("\n".join(md))


def main():
    os.makedirs(OUT, exist_ok=True)
    ms, n = launches_summary()
    t = traffic()
    config5()
    for src, dst in (("r2e_bench_n1.json", "bench_21M_64k.json"), ("r2e_bench_ref.json", "bench_21M_64k_reference_arm.json"),
                     ("r2e_prober_bench_16k.json", "prober_bench_16k.json")):
        if os.path.exists(os.path.join(G, src)):
            line = open(os.path.join(G, src)).read().strip().splitlines()[-1]
            json.dump(json.loads(line), open(os.path.join(OUT, dst), "w"), indent=1)
    print("scoring launches", n, "ms", ms, "traffic", t["dram_bytes_per_launch"], t["limiter"])


if __name__ == "__main__":
    main()
