"""Where does the end-to-end step spend its time?  replay alone / matcher alone / both, pinned host buffers."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
sfe = importlib.import_module("slam-robot_b200"); synth = importlib.import_module("slam-robot_b200.synth")
B = 128
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
A, Bf, pts, q, t = bench.make_inputs(torch, synth, B, dev, 1)
n = B * bench.NFEAT
hA, hB = A.cpu().pin_memory(), Bf.cpu().pin_memory()
h_pts = fe.pinned((n, 2), np.float32); h_pts[...] = pts
h_q, h_t = fe.pinned(q.shape, np.uint32), fe.pinned(t.shape, np.uint32); h_q[...] = q; h_t[...] = t
out = dict(to_xy=fe.pinned((n, 2), np.float32), back_xy=fe.pinned((n, 2), np.float32), status_fwd=fe.pinned((n,), np.int32),
           status_bwd=fe.pinned((n,), np.int32), accepted=fe.pinned((n,), np.uint8), steps=fe.pinned((n,), np.int32))
h_ham = (fe.pinned((n, 2), np.int32), fe.pinned((n, 2), np.int32), fe.pinned((n,), np.uint8))
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def replay(c): return lambda: fe.replay_pairs(hA, hB, h_pts, h_pts, bench.DEPTH, bench.LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT, out=out, chunk_pairs=c)
def ham(): fe.match_hamming256_async(h_q, h_t, h_ham, *bench.RATIO, batch=B); fe.sync()
def both(c):
    def f():
        fe.match_hamming256_async(h_q, h_t, h_ham, *bench.RATIO, batch=B); replay(c)()
    return f
def h2d():
    d = torch.empty_like(A); d.copy_(hA, non_blocking=True); d.copy_(hB, non_blocking=True)
print("H2D of both frame sets alone: %.2f ms" % timeit(h2d))
print("matcher alone (async + sync): %.2f ms" % timeit(ham))
for c in (8, 16, 32, 64):
    print("chunk %3d: replay alone %.2f ms, with matcher %.2f ms" % (c, timeit(replay(c)), timeit(both(c))))
