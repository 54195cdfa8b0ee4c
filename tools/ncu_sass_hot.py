"""Per-opcode / per-line digest of an `ncu --page source --csv --print-source sass` export.

    ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:NAME --launch-count 1 > /tmp/k.csv
    python tools/ncu_sass_hot.py /tmp/k.csv [top]
Prints: executed warp instructions and average active lanes per opcode, stall samples per opcode, and the hottest SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
col = {n: i for i, n in enumerate(hdr)}
S, I, T, W = col["Source"], col["Instructions Executed"], col["Thread Instructions Executed"], col["Warp Stall Sampling (All Samples)"]
ops = collections.defaultdict(lambda: [0, 0, 0])
lines = []
seen = set()
for r in rows[h + 1:]:
    if len(r) <= max(S, I, T, W) or r[0] in seen:       # (the export lists every address once per view: keep the first)
        continue
    seen.add(r[0])
    src = r[S].strip()
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
    op = op.split(".")[0].rstrip(";")
    try:
        i, t, w = int(r[I]), int(r[T]), int(r[W])
    except ValueError:
        continue
    o = ops[op]
    o[0] += i; o[1] += t; o[2] += w
    lines.append((w, i, t, src))
ti = sum(o[0] for o in ops.values()); tw = sum(o[2] for o in ops.values()); tt = sum(o[1] for o in ops.values())
print("total warp-instructions %d, avg active lanes %.2f, stall samples %d" % (ti, tt / max(ti, 1), tw))
print("%-12s %14s %7s %7s %9s %7s" % ("opcode", "warp-inst", "share", "lanes", "samples", "share"))
for op, o in sorted(ops.items(), key=lambda x: -x[1][0])[:top]:
    print("%-12s %14d %7.3f %7.2f %9d %7.3f" % (op, o[0], o[0] / ti, o[1] / max(o[0], 1), o[2], o[2] / max(tw, 1)))
print("\nhottest lines by stall samples")
for w, i, t, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print("%7d samples %12d inst %6.2f lanes  %s" % (w, i, t / max(i, 1), src[:110]))
