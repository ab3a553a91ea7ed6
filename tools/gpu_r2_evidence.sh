#!/bin/bash
# Round-2 evidence run (one GPU): default bench line, ncu launch list of a bench step, ncu --set full captures of the
# scoring kernel (large-batch variant, one mid-step launch; small-batch variant at 64 queries) and of the prober GEMMs,
# and the BASELINE config-5 sweep (batch x depth x round).
mkdir -p gpurun_out
T=${1:-all}
if [ $T = all ] || [ $T = bench ]; then
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2e_bench_n1.json
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2e_bench_ref.json
fi
if [ $T = all ] || [ $T = ncu ]; then
echo "== ncu launch list (our kernels only)"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|prober_|topk" -c 2000 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2e_launches_bench.csv
echo "== ncu full: scoring kernel, large batch"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 85 -c 1 -o gpurun_out/r2e_prof_lean_big python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2e_ncu_lean_big.log 2>&1; echo "rc=$?"
echo "== ncu full: scoring kernel, 64 queries"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 3 -c 1 -o gpurun_out/r2e_prof_lean_b64 python tools/latency.py --batches 64 --k 10 --reps 2 > gpurun_out/r2e_ncu_lean_b64.log 2>&1; echo "rc=$?"
echo "== ncu full: prober GEMMs"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:prober_gemm -s 4 -c 2 -o gpurun_out/r2e_prof_prober python tools/bench_prober.py --rows 16384 --out /dev/null > gpurun_out/r2e_ncu_prober.log 2>&1; echo "rc=$?"
fi
if [ $T = all ] || [ $T = sweep ]; then
echo "== config 5 sweep"; rm -f gpurun_out/r2e_config5_sweep.jsonl
timeout 900 python tools/latency.py --batches 1,8,64,512,4096,65536 --k 1,5,10,50,100 --reps 10 --graph 2>/dev/null >> gpurun_out/r2e_config5_sweep.jsonl
timeout 900 python tools/latency.py --batches 1,8,64,512,4096 --k 1,10,100 --reps 3 --kind later 2>/dev/null >> gpurun_out/r2e_config5_sweep.jsonl
wc -l gpurun_out/r2e_config5_sweep.jsonl
echo "== prober bench"; timeout 300 python tools/bench_prober.py --rows 16384 --with-bm25 --out gpurun_out/r2e_prober_bench_16k.json 2>&1 | tail -n 1 | cut -c1-300
fi
