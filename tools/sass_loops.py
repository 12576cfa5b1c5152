"""List the loops of one kernel in a cubin / .so (by SASS backward branches) with instruction mix, to
check what sits inside the hot loops (spills, conversions, FFMA2, atomics).
usage: python tools/sass_loops.py <lib> <kernel-name-substring> [min_len]"""
import re
import subprocess
import sys
from collections import Counter


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        print(f"== {name}: {len(ins)} instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and tgt in addr and i - addr[tgt] >= min_len:
                    loops.append((addr[tgt], i))
        # keep innermost loops only
        inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
        for lo, hi in inner:
            body = ins[lo:hi + 1]
            c = Counter()
            for _, t in body:
                op = t.split()[1] if t.startswith("@") else t.split()[0]
                c[op.split(".")[0]] += 1
            keys = ["FFMA2", "FFMA", "DFMA", "DADD", "DMUL", "DSETP", "F2F", "I2F", "LDS", "LDL", "STL", "ATOMS", "LOP3", "SEL", "FMNMX", "FMNMX3", "PRMT", "CALL", "STG", "LDG", "MUFU"]
            print(f"  loop {ins[lo][0]:#x}-{ins[hi][0]:#x}: {len(body)} instr  " + " ".join(f"{k}={c[k]}" for k in keys if c[k]))


if __name__ == "__main__":
    main()
