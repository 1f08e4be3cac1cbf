"""CPU-side tests (no GPU): the oracle against the committed cv2 golden vectors (and against a live
cv2 when importable), the C ABI surface of libslamfe.so, the no-CPU-fallback rule, and host logic."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bits_equal

GOLD = os.path.join(ROOT, "tests", "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


# ------------------------------------------------------------------ oracle vs golden (cv2-generated)

def test_mask_properties(po):
    m = po.mask13()
    assert m.shape == (169,) and abs(float(m.astype(np.float64).sum()) - 169.0) < 1e-3
    m2 = m.reshape(13, 13)
    assert m2.argmax() in (6 * 13 + 6, 6 * 13 + 7, 7 * 13 + 6, 7 * 13 + 7)  # centre is at 6.5 (hessian.h:15-16)
    assert np.array_equal(m2, m2.T)


def test_pyramid_hessian_vs_cv2_golden(po):
    """hessian.h:95-126: every level bit-identical to the cv2-built pyramid, scalar-tail columns included (131 = 16*8+3
    columns at level 0, 66 / 33 / 17 above: every tail rule of oracle.h is exercised)."""
    g = gold("pyramid.npz")
    for key, frame in (("hes_a", g["A"]), ("hes_b", g["B"])):
        p = po.Pyramid(frame, 4, po.FLAVOR_HESSIAN)
        for l in range(4):
            assert_bits_equal(p.plane(l), g["%s%d" % (key, l)], "%s level %d" % (key, l))


def test_pyramid_klt_and_brute_vs_cv2_golden(po):
    g = gold("pyramid.npz")
    pk = po.Pyramid(g["A"], 3, po.FLAVOR_KLT)
    for l in range(3):
        for k in range(3):
            assert_bits_equal(pk.plane(l, k), g["klt_a%d_%d" % (l, k)], "klt level %d plane %d" % (l, k))
    pb = po.Pyramid(g["A"], 3, po.FLAVOR_BRUTE)
    for l in range(3):
        assert_bits_equal(pb.plane(l), g["bru_a%d" % l], "brute level %d" % l)


def sha(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("name", ["vga", "hd"])
def test_pyramid_measured_shapes_vs_cv2_digests(po, synth, name):
    """The shapes bench.py measures (C2: 640x480, C3: 1920x1080 / 8 levels): every plane of the oracle's pyramid has
    the SHA-256 of the cv2-built one (tests/golden/make_golden.py; the planes are too large to commit), and the
    border / scalar-tail columns are compared in full so that a mismatch names a column."""
    g = gold("pyramid640.npz")
    seed, h, w, depth = (int(v) for v in g[name + "_seed"])
    fr = synth.make_frames(seed, 1, h, w).numpy()[0]
    assert int(po.gray_u8(fr).astype(np.int64).sum()) == int(g[name + "_gray_sum"]), "synthetic frame generator changed"
    p = po.Pyramid(fr, depth, po.FLAVOR_HESSIAN)
    for l in range(depth):
        got = p.plane(l)
        assert_bits_equal(np.concatenate([got[:, :1], got[:, -8:]], axis=1), g["%s_hes%d_tail" % (name, l)], "level %d tails" % l)
        assert np.array_equal(sha(got), g["%s_hes%d_sha" % (name, l)]), "level %d differs from cv2" % l
    if name == "vga":
        pk, pb = po.Pyramid(fr, 3, po.FLAVOR_KLT), po.Pyramid(fr, 3, po.FLAVOR_BRUTE)
        for l in range(3):
            for k in range(3):
                assert np.array_equal(sha(pk.plane(l, k)), g["vga_klt%d_%d_sha" % (l, k)]), (l, k)
            assert np.array_equal(sha(pb.plane(l)), g["vga_bru%d_sha" % l]), l


def golden_pyr(po, key, depth):
    g = gold("pyramid.npz")
    return po.Pyramid.from_planes([g["%s%d" % (key, l)] for l in range(depth)])


def test_get_patch_bit_exact_vs_cv2_golden(po):
    g = gold("patches.npz")
    p = golden_pyr(po, "hes_a", 4)
    for i in range(len(g["xy"])):
        d, m, q = po.hes_get_patch(p, int(g["level"][i]), g["xy"][i, 0], g["xy"][i, 1])
        assert_bits_equal(d, g["patch"][i], "patch %d at %s level %d" % (i, g["xy"][i], g["level"][i]))
        # statistics use the declared lane/tree order; tier-0 emulates fmaf through float64 (exact)
        assert_bits_equal(np.float32([m, q]), np.float32([g["mean"][i], g["sumsq"][i]]), "stats %d" % i)


def test_brute_hessian_bit_exact_vs_golden(po):
    g = gold("hessian.npz")
    pa, pb = golden_pyr(po, "hes_a", 4), golden_pyr(po, "hes_b", 4)
    for i, (x, y) in enumerate(g["xy"]):
        d, m, q = po.hes_get_patch(pa, 0, x, y)
        s0, d6 = po.hes_brute_hessian(pb, 0, d, m, q, np.float32(x + 0.7), np.float32(y - 0.4))
        assert_bits_equal(np.concatenate([[s0], d6]).astype(np.float32), g["out7"][i], "BruteHessian %d" % i)


@pytest.mark.parametrize("levels", [3, 4])
def test_tracks_bit_exact_on_cv2_planes(po, levels):
    """matcher.cpp:173-206 end to end: tier-0 (cv2.getRectSubPix + NumPy) and the C oracle agree bit for
    bit when they track on the same (cv2-built) planes."""
    g = gold("tracks.npz")
    pa, pb = golden_pyr(po, "hes_a", 4), golden_pyr(po, "hes_b", 4)
    r = po.hes_track_fb(pa, pb, g["xy"], g["xy"], levels)
    pre = "l%d_" % levels
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(r[k], g[pre + k]), k
    assert_bits_equal(r["to_xy"], g[pre + "to_xy"], "to_xy")
    assert_bits_equal(r["back_xy"], g[pre + "back_xy"], "back_xy")
    assert r["newton_steps"] == int(g[pre + "newton_steps"])
    assert r["accepted"].sum() >= 60


def test_tracks_own_pyramid_equal_cv2_pipeline(po):
    """Whole pipeline (oracle pyramid + oracle tracker) against the whole cv2 pipeline (cv2 pyramid + tier-0 tracker):
    0 of 80 features differ -- positions bit for bit, statuses and accept flags identical (north_star asks for 1e-3 px
    and identical lost-track flags)."""
    g, gp = gold("tracks.npz"), gold("pyramid.npz")
    pa, pb = po.Pyramid(gp["A"], 4), po.Pyramid(gp["B"], 4)
    for levels in (3, 4):
        r = po.hes_track_fb(pa, pb, g["xy"], g["xy"], levels)
        pre = "l%d_" % levels
        for k in ("status_fwd", "status_bwd", "accepted"):
            assert np.array_equal(r[k], g[pre + k]), k
        assert_bits_equal(r["to_xy"], g[pre + "to_xy"], "to_xy")
        assert_bits_equal(r["back_xy"], g[pre + "back_xy"], "back_xy")


def test_klt_systems_and_tracks_vs_tier0_golden(po):
    """klt.h on the real cv2 planes (tier-0) -- every Newton iteration's point, its 24 numbers
    A,B,C,RS,VW,U,e,d (klt.h:286-343) and six finite differences, and the forward/backward results -- reproduced bit
    for bit by the C oracle on ITS OWN pyramids (which equal cv2's, test above)."""
    g, gp = gold("klt_tracks.npz"), gold("pyramid.npz")
    pa, pb = po.Pyramid(gp["A"], 3, po.FLAVOR_KLT), po.Pyramid(gp["B"], 3, po.FLAVOR_KLT)
    pts = g["xy"]
    r = po.klt_track_fb(pa, pb, pts, pts)
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(r[k], g[k]), k
    assert_bits_equal(r["to_xy"], g["to_xy"], "klt to_xy")
    assert_bits_equal(r["back_xy"], g["back_xy"], "klt back_xy")
    assert len(g["it_feature"]) > 100
    for j in range(len(g["it_feature"])):
        i, lvl = int(g["it_feature"][j]), int(g["it_level"][j])
        sc = np.float32(1. / (1 << lvl))
        got = po.klt_system(pa, np.float32(pts[i, 0] * sc), np.float32(pts[i, 1] * sc), pb, lvl, g["it_xy"][j, 0], g["it_xy"][j, 1])
        assert_bits_equal(got, g["it_sys24"][j], "klt system of iteration %d (feature %d level %d)" % (j, i, lvl))


def test_brute_search_best_vs_tier0_golden(po):
    """brute.h SearchBest arg-mins from tier-0 (cv2.getRectSubPix): the four cheap passes on 8
    features, and the schedule AS WRITTEN incl. (8, 0.01) = 1600 x 1600 positions (brute.h:158) on 2 features."""
    g, gp = gold("brute_tracks.npz"), gold("pyramid.npz")
    pa, pb = po.Pyramid(gp["A"], 3, po.FLAVOR_BRUTE), po.Pyramid(gp["B"], 3, po.FLAVOR_BRUTE)
    pts = g["xy"]
    r = po.brute_track(pa, pb, pts, pts, fine=po.BRUTE_FINE_FAST)
    assert np.array_equal(r["status"], g["fast_status"])
    assert_bits_equal(r["to_xy"], g["fast_final"][:, :2], "fast schedule positions")
    assert_bits_equal(r["best_sad"], g["fast_final"][:, 2], "fast schedule scores")
    assert r["positions"] == 8 * int(g["fast_positions_per_feature"])
    # every pass on its own: restart the search from the previous pass's arg-min (level-0 passes of feature 0)
    r2 = po.brute_track(pa, pb, pts[:2], pts[:2])   # defaults = the reference's schedule
    assert np.array_equal(r2["status"], g["ref_status"])
    assert_bits_equal(r2["to_xy"], g["ref_final"][:, :2], "reference schedule positions")
    assert_bits_equal(r2["best_sad"], g["ref_final"][:, 2], "reference schedule scores")
    assert r2["positions"] == 2 * int(g["ref_positions_per_feature"]) == 2 * 2560631


def test_hamming_vs_cv2_bfmatcher_golden(po):
    g = gold("hamming.npz")
    idx, dist, ok = po.hamming256_top2(g["q"], g["t"], 4, 5, 64)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert (dist[:, 0] == dist[:, 1]).sum() > 50  # the tie rule is exercised
    exp_ok = (dist[:, 0] <= 64) & (dist[:, 0].astype(np.int64) * 5 < dist[:, 1].astype(np.int64) * 4)
    assert np.array_equal(ok.astype(bool), exp_ok)


def test_good_features_vs_cv2_golden(po, synth):
    """Corner seeding (matcher.cpp:313 + :123-130): the oracle's cornerMinEigenVal map equals OpenCV's bit for bit and
    its goodFeaturesToTrack lists are identical (order included) for four parameter sets on three frames."""
    g = np.load(os.path.join(GOLD, "corners.npz"))
    params = g["params"]
    for name in ("s", "m", "vga"):
        if name == "vga":
            seed, h, w = (int(v) for v in g["vga_seed"])
            fr = synth.make_frames(seed, 1, h, w).numpy()[0]
            assert int(po.gray_u8(fr).astype(np.int64).sum()) == int(g["vga_gray_sum"]), "synthetic frame generator changed"
        else:
            fr = g[name + "_frame"]
        for k, (maxc, q, mind) in enumerate(params):
            c, eig, _ = po.good_features(fr, int(maxc), float(q), float(mind), want_eig=True)
            ref = g["%s_corners%d" % (name, k)]
            assert c.shape == ref.shape and np.array_equal(c, ref), (name, k, len(c), len(ref))
        if name == "vga":
            assert np.array_equal(eig[::40].view(np.uint32), g["vga_eig_rows"].view(np.uint32))
        else:
            assert np.array_equal(eig.view(np.uint32), g[name + "_eig"].view(np.uint32))
    # degenerate inputs: a flat image has no corners; a tiny one does not crash
    assert len(po.good_features(np.full((32, 48, 3), 90, np.uint8), 50, 0.01, 5.0)) == 0
    assert po.good_features(np.zeros((16, 16, 3), np.uint8), 5, 0.01, 1.0).shape == (0, 2)


def test_seed_and_yuyv_oracles_against_independent_restatements(po):
    """8f ranks 3 and 4 have no OpenCV inside: the oracle follows project.h:11-54 / video.cpp:187-223 line by line
    (Eigen itself is absent here, so the quaternion product's operation order is stated from Eigen 3.2 and
    "parity unpinned" to the last bit); checked against independent numpy formulations."""
    rng = np.random.default_rng(4)
    n = 2000
    axis = np.float64([0.1, 0.7, -0.2]); axis /= np.linalg.norm(axis)
    ang = -0.4
    rot = np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]])
    trans = np.float64([0.2, 0.1, -0.3])
    k = np.float64([-0.1, 0.02, 0.001, 300.0, 301.0, 320.0, 240.0])
    pts = np.concatenate([rng.normal(0, 2, (n, 2)), rng.uniform(1, 10, (n, 1)), rng.uniform(0.5, 2, (n, 1))], 1)
    from_xy = rng.uniform(0, 400, (n, 2)).astype(np.float32)
    seed, lv, go = po.seed_features(pts, np.full(n, 5.0), rot, trans, k, from_xy, 640, 480)
    # rotation matrix form of the same projection
    x, y, z, w = rot
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    p = (R @ (pts[:, :3] - trans * pts[:, 3:4]).T).T
    xp, yp = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
    r2 = xp * xp + yp * yp
    d = 1 + r2 * (k[0] + r2 * (k[1] + r2 * k[2]))
    ref = np.stack([xp * d * k[3] + k[5], yp * d * k[4] + k[6]], 1)
    ok = p[:, 2] >= 0.001 * pts[:, 3]
    assert ok.sum() > n // 2 and (~ok).sum() > 0          # both branches of project.h:27 are exercised
    assert np.allclose(seed[ok], ref[ok].astype(np.float32), rtol=1e-5, atol=1e-3)
    assert np.array_equal(seed[~ok], from_xy[~ok])        # behind the lens: the seed stays from_pt
    assert (lv == 3).all()
    assert np.array_equal(go.astype(bool), ~((seed[:, 0] < 0) | (seed[:, 1] < 0) | (seed[:, 0] >= 640) | (seed[:, 1] > 480)))
    # uncertainty >= 100: the seed stays from_pt, levels 6 above 100
    s2, l2, _ = po.seed_features(pts, np.where(np.arange(n) % 2, 100.0, 1e8), rot, trans, k, from_xy, 640, 480)
    assert np.array_equal(s2, from_xy) and set(l2[::2]) == {6} and set(l2[1::2]) == {3}
    # YUYV: vectorised numpy restatement
    buf = rng.integers(0, 256, 4 * 5000, dtype=np.uint8)
    q = buf.reshape(-1, 4).astype(np.int64)
    u, v = q[:, 1] - 128, q[:, 3] - 128
    cb, cr, cg = (u * 454) >> 8, (v * 359) >> 8, (u * 88 + v * 183) >> 8
    px = np.stack([q[:, 0] + cb, q[:, 0] - cg, q[:, 0] + cr, q[:, 2] + cb, q[:, 2] - cg, q[:, 2] + cr], 1).clip(0, 255).astype(np.uint8)
    assert np.array_equal(po.yuyv_to_bgr(buf), px.ravel())


def test_replay_source_reads_the_reference_png_format(tmp_path):
    """ImageSourceFiles (video.h:24-38) reads "<dir>/%08d.png" through cv::imread.  The C++ mirror
    (host/replay_source.hpp, own PNG decoder) must return the bytes cv2.imread returns: colour frames at every zlib
    level (stored / fixed / dynamic Huffman blocks, all five scanline filters), gray, RGBA and palette files; a missing
    frame ends the sequence like main.cpp:518-519."""
    import subprocess
    import cv2
    exe = str(tmp_path / "test_replay_source")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_replay_source.cpp")])
    d = tmp_path / "frames"
    d.mkdir()
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:60, 0:83]
    smooth = np.stack([(xx * 3) % 256, (yy * 4) % 256, ((xx + yy) * 2) % 256], 2).astype(np.uint8)
    frames = []
    for i in range(7):
        img = smooth.copy() if i % 2 else rng.integers(0, 256, (60, 83, 3), dtype=np.uint8)
        img[10:20, 5:40] = 17 * i          # flat runs -> long matches
        frames.append(img)
    cv2.imwrite(str(d / "00000000.png"), frames[0], [cv2.IMWRITE_PNG_COMPRESSION, 0])       # stored blocks
    cv2.imwrite(str(d / "00000001.png"), frames[1], [cv2.IMWRITE_PNG_COMPRESSION, 9])
    cv2.imwrite(str(d / "00000002.png"), frames[2], [cv2.IMWRITE_PNG_COMPRESSION, 1, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FIXED])
    cv2.imwrite(str(d / "00000003.png"), frames[3][..., 0])                                # 8-bit gray
    cv2.imwrite(str(d / "00000004.png"), np.dstack([frames[4], np.full((60, 83), 200, np.uint8)]))  # RGBA
    cv2.imwrite(str(d / "00000005.png"), frames[5], [cv2.IMWRITE_PNG_COMPRESSION, 6, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FILTERED])
    cv2.imwrite(str(d / "00000006.png"), frames[6])
    r = subprocess.run([exe, str(d), "0", "8", str(tmp_path / "out")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0
    for i in range(8):
        raw = open(tmp_path / ("out%d.bin" % i), "rb").read()
        head, _, body = raw.partition(b"\n")
        w, h = (int(v) for v in head.split())
        ref = cv2.imread(str(d / ("%08d.png" % i)), cv2.IMREAD_COLOR)
        if ref is None:
            assert (w, h) == (0, 0) and i == 7
            continue
        assert (h, w) == ref.shape[:2], i
        assert np.array_equal(np.frombuffer(body, np.uint8).reshape(h, w, 3), ref), "frame %d" % i
    # frames 0..6 exist: pairs (0,2) (1,3) (2,4) (3,5) (4,6); (5,7) has no second frame
    assert r.stdout.split()[:4] == ["pairs", "5", "83", "60"]
    # the same frames through LoadSequence (one buffer for sfe_replay_sequence): 7 frames, consistent with the pairs
    assert r.stdout.split()[-3:] == ["sequence", "7", "1"]


def test_hamming_edge_cases(po):
    q = np.zeros((3, 8), np.uint32)
    idx, dist, ok = po.hamming256_top2(q, np.zeros((0, 8), np.uint32))
    assert (idx == -1).all() and (dist == 257).all() and not ok.any()
    t = np.zeros((1, 8), np.uint32)
    t[0, 3] = 0xFF
    idx, dist, ok = po.hamming256_top2(q, t, 4, 5, 256)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all() and (dist[:, 0] == 8).all() and ok.all()
    full = np.full((2, 8), 0xFFFFFFFF, np.uint32)
    _, dist, _ = po.hamming256_top2(q, full)
    assert (dist == 256).all()


# ------------------------------------------------------------------ live cv2 (when importable)

def test_oracle_vs_live_cv2(po):
    t0 = importlib.import_module("oracle.tier0_cv2")
    if not t0.have_cv2():
        pytest.skip("cv2 not importable")
    import cv2
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    assert np.array_equal(po.gray_u8(bgr), cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY))
    img = rng.random((64, 96), dtype=np.float32)
    for sigma in (1.1, 0.8, 0.6):
        assert_bits_equal(po.gauss5(img, sigma), cv2.GaussianBlur(img, (5, 5), sigma, sigmaY=sigma), "blur %.1f" % sigma)
    assert_bits_equal(po.pyrdown(img)[:, 1:-4], cv2.pyrDown(img)[:, 1:-4], "pyrDown interior")
    for _ in range(300):
        n, m = int(rng.integers(6, 14)), int(rng.integers(6, 14))
        cx, cy = float(np.float32(rng.uniform(-3, 99))), float(np.float32(rng.uniform(-3, 67)))
        assert_bits_equal(po.rect_subpix(img, n, m, cx, cy), cv2.getRectSubPix(img, (n, m), (cx, cy)), "subpix")
    q = rng.integers(0, 2 ** 32, (200, 8), dtype=np.uint64).astype(np.uint32)
    t = rng.integers(0, 2 ** 32, (300, 8), dtype=np.uint64).astype(np.uint32)
    ci, cd = t0.hamming_knn2_cv2(q.view(np.uint8), t.view(np.uint8))
    oi, od, _ = po.hamming256_top2(q, t)
    assert np.array_equal(ci, oi) and np.array_equal(cd, od)


# ------------------------------------------------------------------ KLT / brute restatements: properties

def test_klt_and_brute_oracle_properties(po, synth):
    H, W = 120, 160
    A, B = synth.make_pairs(8, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    pts = synth.make_features(2, 40, H, W, margin=20)
    ka, kb = po.Pyramid(A, 3, po.FLAVOR_KLT), po.Pyramid(B, 3, po.FLAVOR_KLT)
    r = po.klt_track_fb(ka, kb, pts, pts)
    truth = synth.true_motion(pts, H, W, 0.004, (1.7, -2.3))
    ok = r["accepted"] == 1
    assert ok.sum() >= 20 and np.median(np.linalg.norm(r["to_xy"] - truth, axis=1)[ok]) < 0.2
    # identical frames: the symmetric-KLT system has A == B == C, RS == VW == 0 and d == 0
    s = po.klt_system(ka, pts[0, 0], pts[0, 1], ka, 0, pts[0, 0], pts[0, 1])
    assert np.allclose(s[0:4], s[4:8], rtol=1e-6) and np.allclose(s[0:4], s[8:12], rtol=1e-6)
    assert np.abs(s[12:16]).max() < 1e-6 and np.abs(s[22:24]).max() < 1e-4
    ba, bb = po.Pyramid(A, 3, po.FLAVOR_BRUTE), po.Pyramid(B, 3, po.FLAVOR_BRUTE)
    rb = po.brute_track(ba, bb, pts, pts, fine=po.BRUTE_FINE_FAST)
    okb = rb["status"] == 0
    assert okb.sum() >= 36 and np.median(np.linalg.norm(rb["to_xy"] - truth, axis=1)[okb]) < 0.2
    # float loop counters (brute.h:105-106): `x += res` accumulates rounding, e.g. (0.2, 0.025) visits 16
    # offsets per axis, not 17
    def count(window, res):
        x, n = np.float32(-window), 0
        while x <= np.float32(window):
            n += 1
            x = np.float32(x + np.float32(res))
        return n * n
    coarse = sum(count(w, r) for w, r in po.BRUTE_COARSE.reshape(-1, 2))
    fine = sum(count(w, r) for w, r in po.BRUTE_FINE_FAST.reshape(-1, 2))
    assert count(0.2, 0.025) == 256 and count(8, 0.01) == 1600 * 1600  # not 1601^2
    assert rb["positions"] == okb.sum() * (2 * coarse + fine)


# ------------------------------------------------------------------ the C ABI surface

def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "slamfe.h")).read()
    return sorted(set(re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(sfe):
    assert declared_symbols() == sorted(sfe.EXPORTS)


def test_library_builds_and_exports_every_symbol(sfe):
    path = sfe.build()
    L = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(L, s), "libslamfe.so does not export %s" % s
    out = subprocess.run(["cuobjdump", "-lelf", path], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out, "library must carry sm_100a code"


@pytest.mark.parametrize("src", ["tests/cpp/test_shim.cpp", "tests/cpp/test_dist.cpp", "tools/replay_dir.cpp"])
def test_cpp_host_layer_compiles_and_links(tmp_path, sfe, src):
    """The reference-side C++ callers (the GpuTracker / MatcherT mirror of host/*.hpp, the multi-GPU example, the replay
    tool) build against include/slamfe.h and link against libslamfe.so without a GPU; the GPU suite runs them."""
    sfe.build()
    csrc = os.path.join(ROOT, "slam-robot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-pthread", "-o", str(tmp_path / "exe"), os.path.join(ROOT, src),
                           "-L" + csrc, "-lslamfe", "-Wl,-rpath," + csrc])


def test_bind_host_to_device_without_gpu_is_a_no_op(sfe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by tests/test_gpu_dist.py on a GPU box")
    before = os.sched_getaffinity(0)
    assert sfe.bind_host_to_device(0) == 0 and os.sched_getaffinity(0) == before


def test_no_cpu_fallback(sfe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfe.SlamFEError):
        sfe.FrontEnd(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "slam-robot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "tier0_cv2" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f


def test_every_header_entry_cites_the_reference():
    txt = open(os.path.join(ROOT, "include", "slamfe.h")).read()
    for ref in ("matcher.cpp:173-206", "hessian.h:95-126", "hessian.h:54-93", "hessian.h:147-172", "klt.h:258-424",
                "brute.h:129-164", "hessian.h:48-52", "matcher.cpp:304"):
        assert ref in txt, ref


# ------------------------------------------------------------------ host logic

def test_synth_is_seeded(synth):
    a1, b1 = synth.make_pairs(3, 2, 48, 64)
    a2, b2 = synth.make_pairs(3, 2, 48, 64)
    assert np.array_equal(a1.numpy(), a2.numpy()) and np.array_equal(b1.numpy(), b2.numpy())
    assert not np.array_equal(a1.numpy(), b1.numpy())
    p = synth.make_features(1, 100, 48, 64, margin=8, border_frac=0.2)
    assert p.shape == (100, 2) and p.min() > 0 and (p[:, 0] < 64).all() and (p[:, 1] < 48).all()
    d = synth.make_descriptors(1, 50, dup_frac=0.2)
    assert d.shape == (50, 8) and d.dtype == np.uint32


def test_shard_ranges(sfe):
    dist = importlib.import_module("slam-robot_b200.dist")
    for n in (0, 1, 7, 65536, 1000003):
        for world in (1, 2, 3, 8):
            r = [dist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == dist.shard_counts(n, world)


def test_bench_reference_arm_runs():
    """bench.py --impl reference prints one JSON line with the contract's keys (tiny sample)."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-pairs", "1"], stdout=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "frame pairs/s"
