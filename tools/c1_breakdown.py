"""Where a live frame's 240 us go: device-side times (CUDA events) of the one-frame pyramid build, the 500-feature tracking
launch and the one-frame corner seeding, beside the wall clock of the host-pointer calls."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sfe = importlib.import_module("slam-robot_b200"); synth = importlib.import_module("slam-robot_b200.synth")
W, H, NF = 640, 480, 500
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
A, B = synth.make_pairs(5, 1, H, W, device=dev)
pts = torch.from_numpy(synth.make_features(9, NF, H, W, margin=16.0).astype(np.float32)).to(dev)
pa = fe.pyramid(W, H, 6, sfe.HESSIAN, 1); pb = fe.pyramid(W, H, 6, sfe.HESSIAN, 1)
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
def ev(): return torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    pa.build(A)
    for levels in (3, 6):
        tp = tt = 0.0
        for it in range(25):
            to = pts.clone()
            e = [ev(), ev(), ev()]
            e[0].record(st); pb.build(B); e[1].record(st)
            r = fe.track_fb(pa, pb, pts, to, levels, 0.001, 10, 0.3, n_per_pair=NF); e[2].record(st)
            torch.cuda.synchronize()
            if it >= 5: tp += e[0].elapsed_time(e[1]); tt += e[1].elapsed_time(e[2])
        print("device side, %d levels: pyramid of one frame %.1f us, tracking of %d features %.1f us (Newton steps/feature %.1f, max %d)" % (
            levels, tp / 20 * 1e3, NF, tt / 20 * 1e3, float(r["steps"].float().mean()), int(r["steps"].max())))
    tg = 0.0
    for it in range(25):
        e = [ev(), ev()]
        e[0].record(st); fe.good_features(B, 120, 0.01, 20.0); e[1].record(st); torch.cuda.synchronize()
        if it >= 5: tg += e[0].elapsed_time(e[1])
    print("device side: corner seeding of one frame %.1f us" % (tg / 20 * 1e3))
