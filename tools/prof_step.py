"""Small fixed workload for ncu captures: a few steps of the bench path at a reduced batch."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
A, Bf, pts, q, t = bench.make_inputs(torch, synth, B, dev, 1)
n = B * bench.NFEAT
from_xy = torch.from_numpy(pts).to(dev); to_xy = torch.empty_like(from_xy)
q_d, t_d = torch.from_numpy(q.view(np.int32)).to(dev), torch.from_numpy(t.view(np.int32)).to(dev)
pa = fe.pyramid(bench.W, bench.H, bench.DEPTH, sfe.HESSIAN, B); pb = fe.pyramid(bench.W, bench.H, bench.DEPTH, sfe.HESSIAN, B)
for _ in range(steps):
    pa.build(A); pb.build(Bf)
    to_xy.copy_(from_xy)
    r = fe.track_fb(pa, pb, from_xy, to_xy, bench.LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=bench.NFEAT)
    m = fe.match_hamming256(q_d, t_d, *bench.RATIO, batch=B)
fe.sync(); torch.cuda.synchronize()
print("ok", int(r["accepted"].sum().item()), int(r["steps"].sum().item()))
