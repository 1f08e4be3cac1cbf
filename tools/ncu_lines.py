"""Attribute executed warp-instructions / stall samples of an ncu SASS dump to CUDA source lines.
usage: python tools/ncu_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c listing> <kernel substring>
The nvdisasm listing supplies the line markers; instructions are matched in address order."""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
# --- ncu per-instruction rows
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ncu = []
for r in rows[2:]:
    if len(r) <= iE: continue
    try: ncu.append((int(r[iE]), int(r[iN]), r[iS].strip()))
    except ValueError: pass
# --- nvdisasm listing: walk the kernel's section
lines = open(dis).read().splitlines()
inside = False; cur = ("?", 0); seq = []
for ln in lines:
    if ln.startswith(".section") or "\t.section" in ln:
        inside = (".text." in ln and kname in ln)
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", ln):
        seq.append(cur)
print("ncu instr rows %d, nvdisasm instr %d" % (len(ncu), len(seq)))
n = min(len(ncu), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (e, s, _), key in zip(ncu[:n], seq[:n]):
    a = agg[key]; a[0] += e; a[1] += s; a[2] += 1
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
srcs = {}
def text(f, l):
    import os
    for root in ("/root/repo/slam-robot_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][l - 1].strip()[:80] if l - 1 < len(srcs[p]) else ""
    return ""
print("%-22s %6s %7s %7s %5s  %s" % ("file:line", "sass", "exec%", "samp%", "", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print("%-22s %6d %6.2f%% %6.2f%%        %s" % ("%s:%d" % key, a[2], 100.0 * a[0] / tot_e, 100.0 * a[1] / max(tot_s, 1), text(*key)))
