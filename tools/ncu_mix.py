"""Instruction mix + stall samples per opcode from `ncu --page source --csv` output."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data)
tsm = sum(int(r[ismp]) for r in data)
ops, smp = collections.Counter(), collections.Counter()
DETAIL = ('F2F', 'I2F', 'F2I', 'FRND', 'MUFU', 'DSETP', 'DMNMX', 'ATOMS', 'LDG', 'LDS', 'STS', 'SHFL', 'STG', 'ATOMG', 'RED')
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
    full = m.group(2) if m else '?'
    op = full.split('.')[0]
    key = full if op in DETAIL else op
    ops[key] += int(r[ia])
    smp[key] += int(r[ismp])
print('total warp instr', tot, 'samples', tsm)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{k:30s} {v:12d} {v / tot * 100:5.1f}%  samples {smp[k] / max(tsm, 1) * 100:5.1f}%")
