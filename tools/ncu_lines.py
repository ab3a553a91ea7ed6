"""Aggregate an .ncu-rep's executed warp instructions and stall samples per CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout
    hdr, fname = None, '?'
    agg, samp, src = collections.Counter(), collections.Counter(), {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            ln = int(r[0])
            ie = int(r[hdr.index('Instructions Executed')])
            s = int(r[hdr.index('# Samples')] or 0)
        except ValueError:
            continue
        key = (fname, ln)
        agg[key] += ie
        samp[key] += s
        src[key] = r[1]
    tot, ts = sum(agg.values()), sum(samp.values())
    print(f"total warp instructions {tot}, samples {ts}")
    byfile = collections.Counter()
    for (f, _), v in agg.items():
        byfile[f] += v
    for f, v in byfile.most_common():
        print(f"  {f:24s} {v / tot * 100:5.1f}%")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{k[0]}:{k[1]:<5d} instr {v / tot * 100:5.2f}%  samples {samp[k] / max(ts, 1) * 100:5.2f}%  {src[k].strip()[:100]}")


if __name__ == '__main__':
    main()
