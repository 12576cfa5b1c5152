"""Markdown summary + per-pixel DRAM traffic of an `ncu --set full` capture exported with
`ncu -i X.ncu-rep --page raw --csv`.  usage: ncu_report.py raw.csv pixels_per_launch out.md out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2])
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stall_keys = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]


def g(r, k):
    return float(r[idx[k]] or 0) if k in idx else float('nan')


lines = ["| # | kernel | grid | time us | DRAM rd MB | DRAM wr MB | B/px | DRAM % | regs | warps/SMSP | issue % | fp64 % | xu % | top stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for n, r in enumerate(data):
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
    tus = g(r, 'gpu__time_duration.sum')
    if units[idx['gpu__time_duration.sum']] in ('ns', 'nsecond'):
        tus /= 1e3
    rd, wr = g(r, 'dram__bytes_read.sum'), g(r, 'dram__bytes_write.sum')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
    rd *= scale.get(units[idx['dram__bytes_read.sum']], 1)
    wr *= scale.get(units[idx['dram__bytes_write.sum']], 1)
    # ncu prints per-row units only in the header row of this export; detect by magnitude
    if rd < 1e4:
        rd *= 1e9 if rd < 10 else 1e6
    if wr < 1e4:
        wr *= 1e9 if wr < 10 else 1e6
    vals = sorted(((g(r, k), k) for k in stall_keys), reverse=True)
    tot = sum(v for v, _ in vals) or 1
    stalls = ', '.join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {v / tot * 100:.0f}%" for v, k in vals[:3])
    bpp = (rd + wr) / npx
    traffic.setdefault(name, []).append(bpp)
    lines.append(f"| {n} | {name[:44]} | {r[idx['Grid Size']]} | {tus:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {bpp:.1f} | "
                 f"{g(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {int(g(r, 'launch__registers_per_thread'))} | "
                 f"{g(r, 'smsp__warps_active.avg.per_cycle_active'):.1f} | {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
                 f"{g(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
                 f"{g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.0f} | {stalls} |")
open(sys.argv[3], 'w').write('\n'.join(lines) + '\n')
json.dump({"pixels_per_launch": npx, "dram_bytes_per_pixel": traffic}, open(sys.argv[4], 'w'), indent=1)
print('\n'.join(lines))
