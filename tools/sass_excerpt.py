"""Write the SASS of one inner loop of a kernel with its instruction mix (how profiles/r0x_sass_*.txt are made).
usage: python tools/sass_excerpt.py <lib> <kernel-name-substring> <out.txt> "<title>" [--has OP[,OP...]] [--lacks OP[,OP...]] [--min N] [--text REGEX]
The loop is the first innermost loop (by backward branch) of at least N instructions whose opcodes include every
--has entry and none of the --lacks entries (prefix match, e.g. STG.E.128, F2F)."""
import argparse
import re
import subprocess
from collections import Counter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib"); ap.add_argument("kernel"); ap.add_argument("out"); ap.add_argument("title")
    ap.add_argument("--has", default=""); ap.add_argument("--lacks", default=""); ap.add_argument("--min", type=int, default=100); ap.add_argument("--text", default="", help="regex some instruction of the loop must match")
    a = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
    cur, ins, name = None, [], None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if ins:
                break
            cur = m.group(1)
            continue
        if cur and a.kernel in cur:
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
            if m:
                name = cur
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {x: i for i, (x, _) in enumerate(ins)}
    loops = []
    for i, (x, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < x and tgt in addr and i - addr[tgt] >= a.min:
                loops.append((addr[tgt], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    op = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0])
    has = [h for h in a.has.split(",") if h]
    lacks = [h for h in a.lacks.split(",") if h]
    for lo, hi in inner:
        ops = [op(t) for _, t in ins[lo:hi + 1]]
        if a.text and not any(re.search(a.text, t) for _, t in ins[lo:hi + 1]):
            continue
        if all(any(o.startswith(h) for o in ops) for h in has) and not any(any(o.startswith(h) for o in ops) for h in lacks):
            c = Counter(o.split(".")[0] for o in ops)
            with open(a.out, "w") as f:
                f.write(f"# {a.title}\n# kernel: {name}\n# cuobjdump -sass {a.lib} (sm_100a), innermost tile loop, {hi - lo + 1} instructions\n")
                f.write("# instruction mix: " + ", ".join(f"{k} {v}" for k, v in c.most_common()) + "\n\n")
                for x, t in ins[lo:hi + 1]:
                    f.write(f"/*{x:05x}*/  {t}\n")
            print(a.out, hi - lo + 1, "instructions")
            return
    raise SystemExit("no loop matches")


if __name__ == "__main__":
    main()
