"""BASELINE config 3 timing: 1920x1080 frames, 5000 features per frame, 8-level pyramid, one B200.
A step = B frame pairs: one build of the 2B pyramids + forward/backward tracking (the bench.py step at config 3)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H, NF, D = 1920, 1080, 5000, 8
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
As, Bs = [], []
for c0 in range(0, B, 8):
    a, b = synth.make_pairs(100 + c0, min(8, B - c0), H, W, device=dev)
    As.append(a); Bs.append(b)
frames = torch.cat(As + Bs).contiguous()      # slots [0, B) first frames, [B, 2B) second frames
pts = np.concatenate([synth.make_features(7 + p, NF, H, W, margin=16.0) for p in range(B)]).astype(np.float32)
from_xy = torch.from_numpy(pts).to(dev); to_xy = from_xy.clone()
pyr = fe.pyramid(W, H, D, sfe.HESSIAN, 2 * B)
def step():
    pyr.build(frames)
    to_xy.copy_(from_xy)
    return fe.track_fb(pyr, pyr, from_xy, to_xy, D, 0.001, 10, 0.3, n_per_pair=NF, from_first=0, to_first=B)
with torch.cuda.stream(st):
    for _ in range(2): r = step()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(st); pyr.build(frames); e[1].record(st)
    for _ in range(3): r = step()
    e[2].record(st)
torch.cuda.synchronize()
ms_p = e[0].elapsed_time(e[1]); ms = e[1].elapsed_time(e[2]) / 3
print("C3 1920x1080, %d features, %d levels, %d pairs/step: %.2f ms/step (pyramids %.2f ms) -> %.0f frame pairs/s, %.2f M features/s; "
      "Newton steps/feature %.1f, accepted %.3f" % (NF, D, B, ms, ms_p, B / (ms * 1e-3), B * NF / (ms * 1e-3) / 1e6,
      float(r["steps"].sum().item()) / (B * NF), float(r["accepted"].float().mean().item())))
