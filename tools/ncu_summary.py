"""Summarise an .ncu-rep (read on the CPU box): key counters + stall breakdown + hottest SASS lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index] [top_n]
"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print(f'--- launch {n}: {r[hdr.index("Kernel Name")][:90]}')
        for w in WANT:
            if w in hdr:
                print(f'  {w:64s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}')
    src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li),
                                           '--launch-count', '1']))))
    hi = [i for i, r in enumerate(src) if r and r[0] == 'Address'][0]
    h = src[hi]
    idx = {}
    for i, name in enumerate(h):
        idx.setdefault(name, i)
    stalls = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]

    def I(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    data, seen = [], set()
    for r in src[hi + 1:]:
        if len(r) < len(h) or r[0] == 'Address' or r[0] in seen:
            continue
        seen.add(r[0])
        data.append(r)
    ns = sum(I(r[idx['# Samples']]) for r in data)
    ni = sum(I(r[idx['Instructions Executed']]) for r in data)
    print(f'\nstall samples {ns}, warp instructions {ni}')
    tot = {s: sum(I(r[idx[s]]) for r in data) for s in stalls}
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
        print(f'  {s:26s} {100 * v / max(ns, 1):5.1f}%')
    print('\nhottest SASS (samples, executions, instruction, top stalls)')
    for r in sorted(data, key=lambda r: -I(r[idx['# Samples']]))[:top]:
        st = sorted(((s, I(r[idx[s]])) for s in stalls), key=lambda x: -x[1])[:2]
        print(f"  {r[idx['# Samples']]:>7s} {r[idx['Instructions Executed']]:>11s}  {r[idx['Source']][:70]:70s} {st}")
    print('\nmost executed SASS')
    for r in sorted(data, key=lambda r: -I(r[idx['Instructions Executed']]))[:12]:
        print(f"  {r[idx['Instructions Executed']]:>11s}  {r[idx['Source']][:80]}")


if __name__ == '__main__':
    main()
