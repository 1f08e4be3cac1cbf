"""Times the dead-header trackers (klt.h, brute.h) on the bench workload: pairs/s for sfe_klt_track_fb / sfe_brute_track."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
A, Bf, pts, q, t = bench.make_inputs(torch, synth, B, dev, 1)
n = B * bench.NFEAT
from_xy = torch.from_numpy(pts).to(dev); to_xy = from_xy.clone()
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
def timed(f, reps=3):
    with torch.cuda.stream(st):
        f(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): f()
        e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for flavor, name in ((sfe.KLT, "klt"), (sfe.HESSIAN, "hessian")):
    pa = fe.pyramid(bench.W, bench.H, bench.DEPTH, flavor, B); pb = fe.pyramid(bench.W, bench.H, bench.DEPTH, flavor, B)
    ms_p = timed(lambda: (pa.build(A), pb.build(Bf)))
    if flavor == sfe.KLT:
        f = lambda: (to_xy.copy_(from_xy), fe.klt_track_fb(pa, pb, from_xy, to_xy, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT))
    else:
        f = lambda: (to_xy.copy_(from_xy), fe.track_fb(pa, pb, from_xy, to_xy, bench.LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT))
    ms_t = timed(f)
    r = f()[1]
    print("%s: pyramids %.3f ms, tracking %.3f ms per %d pairs -> %.0f pairs/s; Newton steps/feature %.2f, accepted %.3f" % (
        name, ms_p, ms_t, B, B / ((ms_p + ms_t) * 1e-3), float(r["steps"].sum().item()) / n, float(r["accepted"].float().mean().item())))
