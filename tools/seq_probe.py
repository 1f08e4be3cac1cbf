"""sfe_replay_sequence against sfe_replay_pairs on the same pairs of a synthetic alternating-camera sequence (pair i =
(frame i, frame i + 2)): end-to-end pairs/s with pinned host buffers.  The sequence entry uploads and builds every
frame once per chunk instead of twice."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
sfe = importlib.import_module("slam-robot_b200"); synth = importlib.import_module("slam-robot_b200.synth")
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = 2
fe = sfe.FrontEnd(0)
A, B, pts, q, t = bench.make_inputs(torch, synth, (NP + S) // 2 + 1, torch.device("cuda", 0), 1)
nfr = NP + S
seq = torch.empty((nfr,) + tuple(A.shape[1:]), dtype=torch.uint8)
# two cameras alternate (main.cpp:503-519): frames 4j, 4j+2 are a related pair, 4j+1, 4j+3 another one
k = 0
for j in range(0, nfr, 4):
    for off, src in ((0, A), (2, B)):
        if j + off < nfr: seq[j + off] = src[k].cpu()
    for off, src in ((1, A), (3, B)):
        if j + off < nfr: seq[j + off] = src[k + 1].cpu() if k + 1 < len(A) else src[k].cpu()
    k += 2
    if k + 1 >= len(A): k = 0
seq = seq.pin_memory()
n = NP * bench.NFEAT
h_pts = fe.pinned((n, 2), np.float32); h_pts[...] = np.resize(pts, (n, 2))   # the same feature lists, cycled
mk = lambda: dict(to_xy=fe.pinned((n, 2), np.float32), back_xy=fe.pinned((n, 2), np.float32), status_fwd=fe.pinned((n,), np.int32),
                  status_bwd=fe.pinned((n,), np.int32), accepted=fe.pinned((n,), np.uint8), steps=fe.pinned((n,), np.int32))
o1, o2 = mk(), mk()
fa, fb = seq[:NP], seq[S:]
def timeit(f, reps=4):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
kw = dict(depth=bench.DEPTH, levels=bench.LEVELS, thr=bench.THR, maxit=bench.MAXIT, fb_max=bench.FB_MAX, n_per_pair=bench.NFEAT)
tp = timeit(lambda: fe.replay_pairs(fa, fb, h_pts, h_pts, out=o1, **kw))
ts = timeit(lambda: fe.replay_sequence(seq, S, h_pts, h_pts, out=o2, **kw))
same = all(np.array_equal(o1[k], o2[k]) for k in ("accepted", "status_fwd", "status_bwd", "steps")) and np.array_equal(o1["to_xy"].view(np.uint32), o2["to_xy"].view(np.uint32))
print("%d pairs of a stride-2 sequence: replay_pairs %.2f ms (%.0f pairs/s, %.0f MB uploaded), replay_sequence %.2f ms (%.0f pairs/s, %.0f MB); "
      "identical results: %s; Newton steps/feature %.1f" % (NP, tp * 1e3, NP / tp, 2 * NP * 0.9216, ts * 1e3, NP / ts,
      (NP + 8 * S * 1.1) * 0.9216, same, float(o2["steps"].mean())))
