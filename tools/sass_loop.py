"""Static view of a kernel's hottest loop: dumps the SASS of one function of libccrs_b200.so, finds the backward branch that
is the smallest one holding >= 80 DFMAs (the observation loop of K2) and prints its opcode histogram.
Usage: python tools/sass_loop.py '<mangled-name substring>' [lib]"""
import collections, re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "camera-intrinsic-calibration-rs_b200", "libccrs_b200.so")
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for fn in funcs[1:]:
    name = fn.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in fn.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)*`?\(?\.?L?_?x?_?\d*\)?\s*$", s)
        t = re.search(r"0x([0-9a-f]+)", s) if "BRA" in s else None
        if t:
            ta = int(t.group(1), 16)
            if ta in addr and addr[ta] < i:
                body = ins[addr[ta]:i + 1]
                nd = sum(1 for _, x in body if re.sub(r"^@!?U?P\d+\s+", "", x).startswith("DFMA"))
                # the smallest loop that holds the Gram DFMAs (>= 80): the observation loop, not an enclosing one
                if nd >= 80 and (best is None or (i - addr[ta]) < (best[2] - best[1])):
                    best = (nd, addr[ta], i)
    print(name)
    if not best:
        print("  no loop found"); continue
    nd, lo, hi = best
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for _, s in ins[lo:hi + 1])
    tot = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"  loop: {tot} instructions ({fp64} FP64, {tot - fp64} other), static count incl. untaken paths")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
