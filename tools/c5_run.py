"""BASELINE config 5 on N GPUs: query rows sharded, train set broadcast from rank 0, top-2 rows all-gathered over NCCL
(slam-robot_b200/dist.py).  Launch: python -m torch.distributed.run --nproc-per-node N tools/c5_run.py [nq] [nt]
Checks the gathered result against a single-GPU run of the full problem on rank 0 and prints the timing."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
sfe = importlib.import_module("slam-robot_b200"); sd = importlib.import_module("slam-robot_b200.dist")
synth = importlib.import_module("slam-robot_b200.synth")
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
fe = sfe.FrontEnd(local)
fe.use_torch_stream()
t_np = synth.make_descriptors(11, nt, dup_frac=0.001)
q_np = synth.make_descriptors(12, nq, dup_frac=0.01, source=t_np)
q = torch.from_numpy(q_np.view(np.int32)).to(dev)
t = torch.from_numpy(t_np.view(np.int32)).to(dev) if rank == 0 else torch.zeros((nt, 8), dtype=torch.int32, device=dev)
match = lambda qr, tr: fe.match_hamming256(qr, tr, 4, 5, 80)
for it in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    idx, dst, ok = sd.match_hamming256_sharded(match, q, t, nq)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    ri, rd, rok = fe.match_hamming256(q, torch.from_numpy(t_np.view(np.int32)).to(dev), 4, 5, 80)
    torch.cuda.synchronize()
    same = bool(torch.equal(idx, ri) and torch.equal(dst, rd) and torch.equal(ok, rok))
    print("C5 %d x %d on %d GPUs: %.3f s (%.2f T comparisons/s incl. broadcast + all-gather), identical to the single-GPU result: %s" % (
        nq, nt, world, dt, nq * nt / dt / 1e12, same))
    assert same
dist.destroy_process_group()
