"""Bring-up / measurement probe for the tensor-core Hamming matcher: parity against the oracle on a few shapes,
then CUDA-event timings of both kernels at the bench shape (512 x 2000 x 2000) and at 64k x 64k."""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
from oracle import pyoracle as po

fe = sfe.FrontEnd(0)
ok_all = True
for (nq, nt, batch) in [(128, 256, 1), (100, 200, 1), (300, 700, 2), (2000, 2000, 2)]:
    t = synth.make_descriptors(1, nt * batch, dup_frac=0.05)
    q = synth.make_descriptors(2, nq * batch, dup_frac=0.3, source=t)
    for impl in (1, 2):
        fe.hamming_impl(impl)
        idx, dist, ok = fe.match_hamming256(q, t, 4, 5, 80, batch=batch)
        bad = 0
        for b in range(batch):
            oi, od, oo = po.hamming256_top2(q[b * nq:(b + 1) * nq], t[b * nt:(b + 1) * nt], 4, 5, 80)
            sl = slice(b * nq, (b + 1) * nq)
            m = (idx[sl] != oi).any(1) | (dist[sl] != od).any(1)
            bad += int(m.sum())
            if m.any() and impl == 2:
                r = np.flatnonzero(m)[:4]
                for i in r:
                    print("   row", i, "got", idx[sl][i], dist[sl][i], "want", oi[i], od[i])
        print("shape", (nq, nt, batch), "impl", impl, "mismatching rows:", bad, flush=True)
        ok_all &= bad == 0
if not ok_all:
    sys.exit(1)

def timeit(nq, nt, batch, impl, reps=5):
    fe.hamming_impl(impl)
    t = torch.from_numpy(synth.make_descriptors(1, nt * batch, dup_frac=0.05).view(np.int32)).cuda()
    q = torch.from_numpy(synth.make_descriptors(2, nq * batch, dup_frac=0.0).view(np.int32)).cuda()
    fe.use_torch_stream()
    out = fe.match_hamming256(q, t, 4, 5, 80, batch=batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fe.match_hamming256(q, t, 4, 5, 80, batch=batch, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("nq %d nt %d batch %d impl %d: %.3f ms  %.1f G comparisons/s" % (nq, nt, batch, impl, ms, nq * nt * batch / ms / 1e6), flush=True)
    return out

for shape in [(2000, 2000, 512), (65536, 65536, 1), (262144, 262144, 1), (500, 500, 1)]:
    a = timeit(*shape, 1)
    b = timeit(*shape, 2)
    print("   identical:", all(bool((x == y).all()) for x, y in zip(a, b)))
fe.hamming_impl(0)
