"""Per-SASS-opcode and hottest-instruction breakdown of the `ncu --page source --csv` dump.
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [n]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
body = []
for r in rows[2:]:
    if len(r) <= iE: continue
    try: e = int(r[iE]); s = int(r[iN])
    except ValueError: continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    op = m.group(1).split(".")[0] if m else "?"
    ops[op] += e; samp[op] += s; tot += e; tots += s
    body.append((e, s, r[iA], r[iS]))
print("total warp-instructions executed: %d, samples %d" % (tot, tots))
print("%-10s %12s %7s %9s" % ("opcode", "executed", "share", "samples%"))
for op, e in ops.most_common(28):
    print("%-10s %12d %6.1f%% %8.1f%%" % (op, e, 100.0 * e / tot, 100.0 * samp[op] / max(tots, 1)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("\nhottest instructions by stall samples:")
for e, s, a, src in sorted(body, key=lambda x: -x[1])[:n]:
    print("%8d samples %10d exec  %s  %s" % (s, e, a, src[:90]))
