"""Times sfe_good_features_dev (CUDA events) on a batch of frames and the oracle on the host; used for DESIGN.md."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
from oracle import pyoracle as po
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
frames = torch.cat([synth.make_frames(i, min(32, B - i), H, W, device=dev) for i in range(0, B, 32)]).contiguous()
with torch.cuda.stream(st):
    for _ in range(3): xy, cnt = fe.good_features(frames, 120, 0.01, 20.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): xy, cnt = fe.good_features(frames, 120, 0.01, 20.0)
    e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
f0 = frames[0].cpu().numpy()
t0 = time.perf_counter(); oc = po.good_features(f0, 120, 0.01, 20.0); cpu = time.perf_counter() - t0
print("B=%d %dx%d: %.3f ms/batch, %.2f us/frame (%.0f frames/s); corners/frame %.1f; oracle (1 thread) %.1f ms/frame" % (
    B, W, H, ms, ms * 1e3 / B, B / (ms * 1e-3), float(cnt.float().mean()), cpu * 1e3))
assert np.array_equal(xy[0, :int(cnt[0])].cpu().numpy(), oc)
