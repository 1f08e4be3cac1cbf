"""Counts, per kernel of the built libslamfe.so, the SASS mnemonics that show which hardware paths the code uses:
tcgen05 (UTCIMMA / LDTM / UTCBAR), the TMA unit (UBLKCP / UTMALDG, SYNCS = mbarrier), cp.async (LDGSTS), packed FP32
(FFMA2 / FMUL2 / FADD2), integer dot products (IDP), POPC, three-input min/max (VIMNMX3).  Runs without a GPU:
    python tools/sass_evidence.py > profiles/sass_mnemonics_r2.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "slam-robot_b200", "csrc", "libslamfe.so")
pat = re.compile(r"^(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTCCP|UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS|FFMA2|FMUL2|FADD2|VIMNMX3|IDP|POPC|HMMA|IMMA)$")
cnt = collections.defaultdict(collections.Counter)
name = None
for l in subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        mm = re.search(r"\d+([a-z0-9_]+_kernel)(.*)", m.group(1))
        name = (mm.group(1) + " " + mm.group(2)[:30]) if mm else m.group(1)[-40:]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and name:
        op = m.group(1).split(".")[0]
        if pat.match(op):
            cnt[name][op] += 1
print("# static SASS mnemonic counts per kernel of libslamfe.so (sm_100a), tools/sass_evidence.py")
for n, c in sorted(cnt.items()):
    print("%-64s %s" % (n, " ".join("%s=%d" % kv for kv in sorted(c.items()))))
