"""The C++ host mirror (slam-robot_b200/host: GpuTracker duck type + Matcher) on the GPU, compared with
the CPU oracle driven the way matcher.cpp drives its tracker."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_shim_matches_oracle(tmp_path, po, synth, sfe):
    sfe.build()
    exe = str(tmp_path / "test_shim")
    csrc = os.path.join(ROOT, "slam-robot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                           "-L" + csrc, "-lslamfe", "-Wl,-rpath," + csrc])
    H, W = 240, 320
    A, B = synth.make_pairs(17, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    corners = synth.make_features(23, 90, H, W, margin=12, border_frac=0.1)
    A.tofile(tmp_path / "A.bin")
    B.tofile(tmp_path / "B.bin")
    corners.tofile(tmp_path / "c.bin")
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, str(tmp_path / "A.bin"), str(tmp_path / "B.bin"), str(tmp_path / "c.bin"), str(W), str(H), str(out)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    lines = open(out).read().splitlines()
    assert lines[0] == "duck_mismatches 0"

    oa, ob = po.Pyramid(A, 6), po.Pyramid(B, 6)
    # 1. batched forward/backward == oracle (3 levels)
    o3 = po.hes_track_fb(oa, ob, corners, corners, 3)
    fb = [l.split() for l in lines if l.startswith("fb ")]
    assert len(fb) == len(corners)
    for i, f in enumerate(fb):
        x, y = np.float32(float.fromhex(f[2])), np.float32(float.fromhex(f[3]))
        assert (x, y) == (o3["to_xy"][i, 0], o3["to_xy"][i, 1]), i
        assert (int(f[4]), int(f[5]), int(f[6])) == (o3["status_fwd"][i], o3["status_bwd"][i], o3["accepted"][i])

    # 2. Matcher: frame A is a keyframe that creates one feature per corner (empty occupancy grid);
    #    frame B tracks them with 6 levels (new points have uncertainty 1e8 > 100, matcher.cpp:227-229)
    a_line = [l for l in lines if l.startswith("after_A")][0].split()
    assert int(a_line[2]) == len(corners) and int(a_line[4]) == 1 and int(a_line[6]) == 1 and int(a_line[8]) == len(corners)
    o6 = po.hes_track_fb(oa, ob, corners, corners, 6)
    exp = [(i, o6["to_xy"][i]) for i in range(len(corners)) if o6["accepted"][i]]
    obs = [l.split() for l in lines if l.startswith("obs ")]
    assert len(obs) == len(exp) and len(exp) > 40
    for (i, xy), o in zip(exp, obs):
        assert int(o[1]) == 1 and int(o[2]) == i
        assert np.float32(float.fromhex(o[3])) == xy[0] and np.float32(float.fromhex(o[4])) == xy[1]
    b_line = [l for l in lines if l.startswith("after_B")][0].split()
    assert int(b_line[6]) == (0 if len(exp) >= 40 else 1)  # >= 40 matches: not a keyframe (matcher.cpp:353)


def test_replay_dir_tool_matches_oracle(tmp_path, po, synth, sfe):
    """The reference's --load replay format end to end on the GPU: PNG frames written with cv2 ("%08d.png", alternating
    cameras -> pairs (id, id+2)) -> tools/replay_dir (PNG decode, sfe_good_features, sfe_replay_pairs) == the oracle."""
    import cv2
    sfe.build()
    exe = str(tmp_path / "replay_dir")
    csrc = os.path.join(ROOT, "slam-robot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-o", exe, os.path.join(ROOT, "tools", "replay_dir.cpp"),
                           "-L" + csrc, "-lslamfe", "-Wl,-rpath," + csrc])
    H, W = 240, 320
    d = tmp_path / "rec"
    d.mkdir()
    # two interleaved cameras: ids 0,2,4 are one moving view, ids 1,3 another
    A, B = synth.make_pairs(51, 3, H, W)
    frames = {0: A[0].numpy(), 2: B[0].numpy(), 4: B[2].numpy(), 1: A[1].numpy(), 3: B[1].numpy()}
    for i, f in frames.items():
        cv2.imwrite(str(d / ("%08d.png" % i)), f)
    out = tmp_path / "tracks.txt"
    r = subprocess.run([exe, str(d), "8", str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("pair ")]
    assert len(lines) == 3                                  # (0,2) (1,3) (2,4); (3,5) has no second frame
    rows = [l.split() for l in open(out)]
    for p, (a, b) in enumerate(((0, 2), (1, 3), (2, 4))):
        corners = po.good_features(frames[a], 120, 0.01, 20.0)
        oa, ob = po.Pyramid(frames[a], 6), po.Pyramid(frames[b], 6)
        o = po.hes_track_fb(oa, ob, corners, corners, 3)
        assert int(lines[p][3]) == len(corners) and int(lines[p][5]) == int(o["accepted"].sum()), (p, lines[p])
        mine = [r_ for r_ in rows if int(r_[0]) == p]
        assert len(mine) == len(corners)
        got = np.float32([[float.fromhex(r_[4]), float.fromhex(r_[5])] for r_ in mine])
        assert np.array_equal(got.view(np.uint32), o["to_xy"].view(np.uint32))
