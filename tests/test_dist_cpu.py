"""world_size-2 gloo test of the multi-GPU host layer (slam-robot_b200/dist.py) on CPU: the sharded
Hamming matcher (query rows sharded, train set broadcast, results all-gathered) and the frame-pair
sharding must reproduce the single-process result.  The per-rank compute is injected; on CPU it is
the oracle (test infrastructure), on the GPUs it is libslamfe (tests/test_gpu_parity.py, bench.py)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nq, nt, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = importlib.import_module("slam-robot_b200.dist")
    synth = importlib.import_module("slam-robot_b200.synth")
    from oracle import pyoracle as po

    t_np = synth.make_descriptors(1, nt, dup_frac=0.05)
    q_np = synth.make_descriptors(2, nq, dup_frac=0.3, source=t_np)
    q = torch.from_numpy(q_np.view(np.int32))
    # only rank 0 owns the real train set; the others start from garbage and receive the broadcast
    t = torch.from_numpy(t_np.view(np.int32)) if rank == 0 else torch.zeros((nt, 8), dtype=torch.int32)

    def match_fn(qr, tr):
        i, d, ok = po.hamming256_top2(qr.numpy().view(np.uint32), tr.numpy().view(np.uint32), 4, 5, 80, nthreads=1)
        return torch.from_numpy(i), torch.from_numpy(d), torch.from_numpy(ok)

    idx, dst, ok = sd.match_hamming256_sharded(match_fn, q, t, nq)
    ri, rd, rok = po.hamming256_top2(q_np, t_np, 4, 5, 80, nthreads=1)
    assert np.array_equal(idx.numpy(), ri) and np.array_equal(dst.numpy(), rd) and np.array_equal(ok.numpy(), rok)

    # frame-pair sharding: every pair is owned by exactly one rank; gathered rows come back in order
    lo, hi = sd.shard_pairs(11, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1) * torch.ones(1, 3)
    allrows = sd.gather_rows(local, 11)
    assert torch.equal(allrows[:, 0], torch.arange(11, dtype=torch.float32))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("nq,nt", [(301, 257), (2, 5)])
def test_sharded_matcher_gloo_world2(tmp_path, nq, nt):
    port = 29500 + (os.getpid() % 2000) + nq % 7
    mp.spawn(_worker, args=(2, port, nq, nt, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_c_abi_shard_range_matches_host_layer(sfe):
    """sfe_shard_range (the C ABI's block rule, pure host arithmetic: callable without a GPU) == dist.shard_range, and
    the blocks tile [0, n) in rank order -- which is what makes an in-place all-gather of the rows correct."""
    import ctypes
    sd = importlib.import_module("slam-robot_b200.dist")
    L = sfe.lib()
    for n in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            end = 0
            for r in range(world):
                lo, hi = ctypes.c_int64(), ctypes.c_int64()
                assert L.sfe_shard_range(n, r, world, ctypes.byref(lo), ctypes.byref(hi)) == 0
                assert (lo.value, hi.value) == sd.shard_range(n, r, world) and lo.value == end
                end = hi.value
            assert end == n
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert L.sfe_shard_range(5, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) != 0   # rank out of range
