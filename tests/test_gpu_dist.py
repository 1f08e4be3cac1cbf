"""The multi-GPU entries of the C ABI (include/slamfe.h "several GPUs of one box") on the GPU box.

world_size 1 always runs (the communicator is created, every collective degenerates); the 2-rank C++ program needs two
visible GPUs (`gpurun --gpus 2`) and is skipped otherwise.  The host-side sharding arithmetic is covered on CPU by
tests/test_dist_cpu.py."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_match_world1_equals_plain(fe, po, synth):
    import torch
    t = synth.make_descriptors(3, 3000, dup_frac=0.05)
    q = synth.make_descriptors(4, 1234, dup_frac=0.3, source=t)
    fe.dist_init(0, 1)
    try:
        idx, dist, ok = fe.match_hamming256_sharded(q, len(q), t, len(t), 0, 4, 5, 80)          # host entry
        oi, od, oo = po.hamming256_top2(q, t, 4, 5, 80)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od) and np.array_equal(ok, oo)
        dev = torch.device("cuda", 0)
        q_d, t_d = torch.from_numpy(q.view(np.int32)).to(dev), torch.from_numpy(t.view(np.int32)).to(dev)
        gi, gd, gp = fe.match_hamming256_sharded(q_d, len(q), t_d, len(t), 0, 4, 5, 80)        # device entry
        fe.sync()
        assert np.array_equal(gi.cpu().numpy(), oi) and np.array_equal(gd.cpu().numpy(), od) and np.array_equal(gp.cpu().numpy(), oo)
        rows = torch.arange(70, dtype=torch.int32, device=dev).reshape(35, 2)
        assert torch.equal(fe.allgather_rows(rows, 35), rows)
        assert fe.shard_range(10, 2, 4) == (6, 8) and fe.shard_range(10, 0, 4) == (0, 3)
    finally:
        fe.dist_shutdown()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_match_cpp_threads(tmp_path, sfe, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    sfe.build()
    exe = str(tmp_path / "test_dist")
    csrc = os.path.join(ROOT, "slam-robot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_dist.cpp"),
                           "-L" + csrc, "-lslamfe", "-Wl,-rpath," + csrc])
    r = subprocess.run([exe, str(world), "30011", "70001"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "sharded_mismatching_ranks 0" in r.stdout, r.stdout


def test_bind_host_to_device_only_narrows(sfe):
    """sfe_bind_host_to_device: the process ends up on a non-empty subset of the CPUs it was allowed before (the CPUs
    next to the GPU), or nothing changes and 0 is returned; the affinity is restored for the rest of the suite."""
    before = os.sched_getaffinity(0)
    try:
        n = sfe.bind_host_to_device(0)
        after = os.sched_getaffinity(0)
        assert after <= before and len(after) > 0
        assert n == 0 and after == before or n == len(after)
    finally:
        os.sched_setaffinity(0, before)
