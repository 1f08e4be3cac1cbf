"""Times sfe_pyr_build_dev alone (CUDA events) for a given batch; used to tune the pyramid kernels."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sfe = importlib.import_module("slam-robot_b200")
synth = importlib.import_module("slam-robot_b200.synth")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H, D = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (640, 480, 4)
dev = torch.device("cuda", 0)
fe = sfe.FrontEnd(0)
st = torch.cuda.Stream(); fe.set_stream(st.cuda_stream)
frames = torch.cat([synth.make_frames(i, min(32, B - i), H, W, device=dev) for i in range(0, B, 32)]).contiguous()
p = fe.pyramid(W, H, D, sfe.HESSIAN, B)
with torch.cuda.stream(st):
    for _ in range(3): p.build(frames)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): p.build(frames)
    e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = B * p.bytes_per_frame() / 1e9
print("chunk_mb=%s B=%d %dx%d d=%d: %.3f ms/build  %.3f us/frame  %.0f GB/s algorithmic (%.1f%% of 6549)" % (
    os.environ.get("SFE_PYR_CHUNK_MB", "default"), B, W, H, D, ms, ms * 1e3 / B, gb / (ms * 1e-3), 100 * gb / (ms * 1e-3) / 6549.1))
