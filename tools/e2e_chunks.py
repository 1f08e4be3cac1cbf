"""Where the end-to-end replay loses against the resident step: sfe_replay_sequence over the bench's 514-frame sequence for a
range of chunk sizes, beside the resident tracking time of the same pairs."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
sfe = importlib.import_module("slam-robot_b200"); synth = importlib.import_module("slam-robot_b200.synth")
B = 512
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
seq = synth.make_sequence(77, B + 2, bench.H, bench.W, stride=2, device=dev).cpu().pin_memory()
pts = np.concatenate([synth.make_features(7919 + p, bench.NFEAT, bench.H, bench.W, margin=16) for p in range(B)]).astype(np.float32)
n = B * bench.NFEAT
h_pts = fe.pinned((n, 2), np.float32); h_pts[...] = pts
out = dict(to_xy=fe.pinned((n, 2), np.float32), back_xy=fe.pinned((n, 2), np.float32), status_fwd=fe.pinned((n,), np.int32),
           status_bwd=fe.pinned((n,), np.int32), accepted=fe.pinned((n,), np.uint8), steps=fe.pinned((n,), np.int32))
kw = dict(depth=bench.DEPTH, levels=bench.LEVELS, thr=bench.THR, maxit=bench.MAXIT, fb_max=bench.FB_MAX, n_per_pair=bench.NFEAT, out=out)
for chunk in [int(a) for a in sys.argv[1:]] or [0, 16, 32, 64, 128, 256, 512]:
    f = lambda: fe.replay_sequence(seq, 2, h_pts, h_pts, chunk_pairs=chunk, **kw)
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print("chunk %4d: %.2f ms per 512 pairs = %.0f pairs/s (Newton steps/feature %.2f)" % (chunk, dt * 1e3, B / dt, float(out["steps"].mean())), flush=True)
# resident: the same pairs, pyramids built once, one tracking launch
frames = seq.to(dev)
pyr = fe.pyramid(bench.W, bench.H, bench.DEPTH, sfe.HESSIAN, B + 2)
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
fx = torch.from_numpy(pts).to(dev); tx = fx.clone()
with torch.cuda.stream(st):
    for _ in range(2):
        pyr.build(frames); tx.copy_(fx)
        fe.track_fb(pyr, pyr, fx, tx, bench.LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT, from_first=0, to_first=2)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(st); pyr.build(frames); e[1].record(st); tx.copy_(fx)
    r = fe.track_fb(pyr, pyr, fx, tx, bench.LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT, from_first=0, to_first=2)
    e[2].record(st)
torch.cuda.synchronize()
print("resident: pyramids of 514 frames %.3f ms, tracking %.3f ms, Newton steps/feature %.2f" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), float(r["steps"].float().mean())))
