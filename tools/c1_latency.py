"""BASELINE config 1 on the GPU: ONE 640x480 frame pair, 500 features, as the live robot runs it (matcher.cpp:317 one
MakePyramid per new frame at depth 6, then FindMatches at 3 levels) -- latency of a frame through the host-pointer C ABI
(frame upload + pyramid + forward/backward tracking + results back), the `replicas only` case of DESIGN.md section 5."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
W, H, NF = 640, 480, int(sys.argv[1]) if len(sys.argv) > 1 else 500
fe = sfe.FrontEnd(0)
A, B = synth.make_pairs(5, 1, H, W)
A, B = A.numpy(), B.numpy()
pts = synth.make_features(9, NF, H, W, margin=16.0).astype(np.float32)
hA, hB = fe.pinned(A.shape, np.uint8), fe.pinned(B.shape, np.uint8)
hA[...] = A; hB[...] = B
pa = fe.pyramid(W, H, 6, sfe.HESSIAN, 1); pb = fe.pyramid(W, H, 6, sfe.HESSIAN, 1)
pa.build(hA)
def frame(levels):
    pb.build(hB)                                   # the new frame's pyramid (host pointer: H2D inside)
    return fe.track_fb(pa, pb, pts, pts.copy(), levels, 0.001, 10, 0.3, n_per_pair=NF)   # host arrays in and out
for levels in (3, 6):
    for _ in range(5): r = frame(levels)
    t0 = time.perf_counter()
    for _ in range(50): r = frame(levels)
    dt = (time.perf_counter() - t0) / 50
    print("C1 on the GPU: 1 frame, %d features, %d levels: %.0f us per frame end to end (%.0f frames/s), accepted %.3f" % (
        NF, levels, dt * 1e6, 1.0 / dt, float(np.mean(r["accepted"]))))
