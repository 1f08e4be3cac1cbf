"""GPU parity tests proper: the CUDA path, called through the C ABI (libslamfe.so), against the
CPU oracle on identical seeded inputs.  Bar: bit-exact for every output (float results are
compared as bit patterns -- the GPU kernels and the oracle spell the same IEEE operations in the
same order), identical status/accept flags.  north_star's own bar (1e-3 px, identical flags) is
checked as well and is implied by the bit-exact one."""
import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def test_mask_matches_oracle(fe, po):
    assert_bits_equal(fe.mask(), po.mask13(), "mask (hessian.h:11-30)")


@pytest.mark.parametrize("shape,depth", [((480, 640), 6), ((97, 131), 5), ((270, 481), 8), ((33, 47), 3), ((100, 200), 5),
                                         ((64, 72), 4), ((250, 1000), 6)])
@pytest.mark.parametrize("flavor", [0, 1, 2])
def test_pyramid_bit_exact(fe, po, synth, shape, depth, flavor):
    H, W = shape
    frames = synth.make_frames(7 + flavor, 2, H, W).numpy()
    gp = fe.make_pyramid(frames, depth, flavor)
    for f in range(2):
        op = po.Pyramid(frames[f], depth, flavor)
        for l in range(depth):
            for plane in range(3 if flavor == 1 else 1):
                assert_bits_equal(gp.plane(l, f, plane), op.plane(l, plane), "flavor %d frame %d level %d plane %d" % (flavor, f, l, plane))


def test_pyramid_1080p_8_levels_bit_exact(fe, po, synth):
    """BASELINE config 3 geometry: 1920x1080, 8 levels (1920 ... 15x9), with a strided host buffer."""
    H, W = 1080, 1920
    frame = synth.make_frames(3, 1, H, W).numpy()[0]
    padded = np.zeros((H, W * 3 + 64), np.uint8)
    padded[:, :W * 3] = frame.reshape(H, W * 3)
    gp = fe.pyramid(W, H, 8, 0, 1)
    rc = fe.L.sfe_pyr_build(fe.h, gp.h, padded.ctypes.data, padded.strides[0], padded.strides[0] * H, 0, 1)
    assert rc == 0
    op = po.Pyramid(frame, 8)
    for l in range(8):
        assert_bits_equal(gp.plane(l), op.plane(l), "1080p level %d" % l)


@pytest.mark.parametrize("shape,depth,nframes", [((480, 640), 4, 3), ((481, 640), 5, 2), ((97, 640), 3, 2), ((720, 1280), 6, 2),
                                                 ((45, 1280), 2, 1), ((1080, 1920), 4, 1), ((541, 1920), 3, 2), ((16, 640), 2, 2), ((17, 640), 2, 1), ((18, 640), 3, 1)])
def test_pyramid_row_kernel_bit_exact(fe, po, synth, shape, depth, nframes):
    """Widths of 640 / 1280 / 1920 columns take the row-CTA kernel (pyramid_stream.cu: one CTA per row band, neighbours
    through shared row buffers, lagged stages): odd heights, bands shorter than the pipeline fill, several bands."""
    H, W = shape
    frames = synth.make_frames(11, nframes, H, W).numpy()
    gp = fe.make_pyramid(frames, depth, 0)
    for f in range(nframes):
        op = po.Pyramid(frames[f], depth, 0)
        for l in range(depth):
            assert_bits_equal(gp.plane(l, f), op.plane(l), "%dx%d frame %d level %d" % (W, H, f, l))


def test_pyramid_build_into_slot_range_bit_exact(fe, po, synth):
    """sfe_pyr_build with first != 0 (what sfe_replay_pairs and the bench's 2B-slot batch rely on): frames written into
    slots [2, 5) of a batch of 6 leave the other slots alone and equal the oracle."""
    H, W = 480, 640
    frames = synth.make_frames(21, 6, H, W).numpy()
    gp = fe.pyramid(W, H, 4, 0, 6)
    gp.build(frames)
    before = [gp.plane(l, 0).copy() for l in range(4)]
    other = synth.make_frames(22, 3, H, W).numpy()
    gp.build(other, first=2)
    for l in range(4):
        assert_bits_equal(gp.plane(l, 0), before[l], "slot 0 untouched, level %d" % l)
        assert_bits_equal(gp.plane(l, 5), po.Pyramid(frames[5], 4, 0).plane(l), "slot 5 untouched, level %d" % l)
        for f in range(3):
            assert_bits_equal(gp.plane(l, 2 + f), po.Pyramid(other[f], 4, 0).plane(l), "slot %d level %d" % (2 + f, l))


def test_pyramid_large_batch_single_band_bit_exact(fe, po, synth):
    """A batch large enough that every frame is ONE row band (the band heuristic only cuts frames when the batch alone
    cannot fill the GPU): first, middle and last frame of 1024 against the oracle."""
    import torch
    H, W, n = 480, 640, 1024
    base = synth.make_frames(5, 32, H, W, device="cuda")
    frames = base.repeat(n // 32, 1, 1, 1).contiguous()
    frames[n // 2] = torch.flip(base[3], dims=[0])
    frames[n - 1] = torch.flip(base[5], dims=[1])
    gp = fe.pyramid(W, H, 4, 0, n)
    gp.build(frames)
    for f in (0, n // 2, n - 1):
        op = po.Pyramid(frames[f].cpu().numpy(), 4, 0)
        for l in range(4):
            assert_bits_equal(gp.plane(l, f), op.plane(l), "frame %d level %d" % (f, l))


def test_pyramid_row_kernel_band_split_invariance(fe, synth):
    """The same 96 frames built in one call (few row bands per frame) and in calls of 8 frames (many short bands per
    frame, different CTA interleaving) must give bit-identical pyramids: the band overlap, the lagged shared-memory
    exchange and its barriers cannot depend on how the frames are cut (a race would show up as a flaky mismatch)."""
    import torch
    H, W, n = 480, 640, 96
    frames = torch.cat([synth.make_frames(31 + i, 32, H, W, device="cuda") for i in range(n // 32)]).contiguous()
    g1 = fe.pyramid(W, H, 4, 0, n)
    g2 = fe.pyramid(W, H, 4, 0, n)
    for rep in range(3):
        g1.build(frames)
        for f0 in range(0, n, 8):
            g2.build(frames[f0:f0 + 8].contiguous(), first=f0)
        for f in range(0, n, 5):
            for l in range(4):
                assert_bits_equal(g1.plane(l, f), g2.plane(l, f), "rep %d frame %d level %d" % (rep, f, l))


def _features(synth, n, H, W, seed=5, border=0.2):
    return synth.make_features(seed, n, H, W, margin=16, border_frac=border)


def test_get_patches_bit_exact(fe, po, pair640, synth):
    A, _ = pair640
    gp = fe.make_pyramid(A, 6)
    op = po.Pyramid(A, 6)
    rng = np.random.default_rng(3)
    for level in range(6):
        w, h = op.size(level)
        n = 300
        xy = np.stack([rng.uniform(0.02, w - 0.02, n), rng.uniform(0.02, h - 0.02, n)], 1).astype(np.float32)
        # force the corners / edges, incl. the top-right region of OpenCV's getRectSubPix quirk
        xy[:8] = [[0.5, 0.5], [w - 0.5, 0.5], [0.5, h - 0.5], [w - 0.5, h - 0.5], [w - 3.3, 0.2], [w - 6.9, 5.7], [6.4, 6.4], [6.6, 6.6]]
        p, m, q = fe.get_patches(gp, level, xy)
        for i in range(n):
            d, mm, qq = po.hes_get_patch(op, level, xy[i, 0], xy[i, 1])
            assert_bits_equal(p[i], d, "patch level %d #%d at %s" % (level, i, xy[i]))
            assert_bits_equal(np.float32([m[i], q[i]]), np.float32([mm, qq]), "stats level %d #%d" % (level, i))


def test_brute_hessian_bit_exact(fe, po, pair640, synth):
    A, B = pair640
    ga, gb = fe.make_pyramid(A, 6), fe.make_pyramid(B, 6)
    oa, ob = po.Pyramid(A, 6), po.Pyramid(B, 6)
    for level in (0, 2, 5):
        w, h = oa.size(level)
        pts = _features(synth, 200, h, w, seed=11 + level, border=0.3) if level < 5 else \
            np.random.default_rng(1).uniform(0.5, [w - 0.5, h - 0.5], (200, 2)).astype(np.float32)
        xy = pts + np.random.default_rng(2).uniform(-1.5, 1.5, pts.shape).astype(np.float32)
        xy = np.clip(xy, 0.02, [w - 0.02, h - 0.02]).astype(np.float32)
        out = fe.brute_hessian(ga, gb, level, pts, xy)
        for i in range(len(pts)):
            d, m, q = po.hes_get_patch(oa, level, pts[i, 0], pts[i, 1])
            s0, d6 = po.hes_brute_hessian(ob, level, d, m, q, xy[i, 0], xy[i, 1])
            assert_bits_equal(out[i], np.concatenate([[s0], d6]).astype(np.float32), "BruteHessian level %d #%d" % (level, i))


@pytest.mark.parametrize("levels", [3, 6, "mixed"])
def test_track_fb_bit_exact(fe, po, pair640, synth, levels):
    A, B = pair640
    H, W = A.shape[:2]
    n = 600
    pts = _features(synth, n, H, W)
    if levels == "mixed":
        lv = np.where(np.arange(n) % 3 == 0, 6, 3).astype(np.int32)
    else:
        lv = levels
    ga, gb = fe.make_pyramid(A, 6), fe.make_pyramid(B, 6)
    oa, ob = po.Pyramid(A, 6), po.Pyramid(B, 6)
    g = fe.track_fb(ga, gb, pts, pts, lv)
    o = po.hes_track_fb(oa, ob, pts, pts, lv)
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(g[k], o[k]), k
    assert_bits_equal(g["to_xy"], o["to_xy"], "to_xy")
    assert_bits_equal(g["back_xy"], o["back_xy"], "back_xy")
    assert int(g["steps"].sum()) == o["newton_steps"]
    assert np.abs(g["to_xy"] - o["to_xy"]).max() <= 1e-3          # north_star tolerance
    assert g["accepted"].sum() > 0.8 * n * 0.8                    # the synthetic pair is trackable
    truth = synth.true_motion(pts, H, W, 0.004, (1.7, -2.3))
    err = np.linalg.norm(g["to_xy"] - truth, axis=1)[g["accepted"] == 1]
    assert np.median(err) < 0.1


def test_track_fb_black_regions_bit_exact(fe, po, synth):
    """Frames with saturated-black blocks: the blurred planes hold exact zeros, so ScorePatchMatch's
    exact-zero skip (hessian.h:134) triggers for candidate and template pixels -- the routes of the CUDA
    tracker that fold the skip away must not be taken there."""
    H, W = 240, 320
    A, B = synth.make_pairs(17, 1, H, W)
    A, B = A[0].numpy().copy(), B[0].numpy().copy()
    rng = np.random.default_rng(4)
    for img, sh in ((A, 0), (B, 2)):
        for (y, x) in ((40, 60), (150, 200), (100, 20), (0, 250), (200, 0)):
            img[y + sh:y + sh + 40, x + sh:x + sh + 48] = 0
    n = 400
    pts = _features(synth, n, H, W, seed=8, border=0.1)
    # half of the features sit on the edges of the black blocks
    edges = np.float32([[60, 40], [108, 80], [200, 150], [248, 190], [20, 100], [68, 140], [250, 0], [298, 40], [0, 200], [48, 239]])
    pts[: n // 2] = (edges[rng.integers(0, len(edges), n // 2)] + rng.uniform(-9, 9, (n // 2, 2))).clip(1, [W - 2, H - 2]).astype(np.float32)
    ga, gb = fe.make_pyramid(A, 4), fe.make_pyramid(B, 4)
    oa, ob = po.Pyramid(A, 4), po.Pyramid(B, 4)
    assert (oa.plane(0) == 0).sum() > 1000 and (oa.plane(2) == 0).sum() > 10   # the zeros are really there
    g = fe.track_fb(ga, gb, pts, pts, 4)
    o = po.hes_track_fb(oa, ob, pts, pts, 4)
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(g[k], o[k]), k
    # NaN positions (all-black patches: 0/0 statistics) must match as bit patterns too
    assert_bits_equal(g["to_xy"], o["to_xy"], "to_xy")
    assert_bits_equal(g["back_xy"], o["back_xy"], "back_xy")
    assert int(g["steps"].sum()) == o["newton_steps"]
    for level in (0, 2):
        s = np.float32(0.5 ** level)
        out = fe.brute_hessian(ga, gb, level, pts[:200] * s, pts[:200] * s + np.float32(0.3))
        for i in range(200):
            d, m, q = po.hes_get_patch(oa, level, pts[i, 0] * s, pts[i, 1] * s)
            s0, d6 = po.hes_brute_hessian(ob, level, d, m, q, pts[i, 0] * s + np.float32(0.3), pts[i, 1] * s + np.float32(0.3))
            assert_bits_equal(out[i], np.concatenate([[s0], d6]).astype(np.float32), "BruteHessian level %d #%d" % (level, i))


def test_track_fb_batched_pairs(fe, po, synth):
    """n_per_pair batching: 3 independent pairs in one call == three single-pair oracle runs."""
    H, W = 240, 320
    A, B = synth.make_pairs(21, 3, H, W)
    A, B = A.numpy(), B.numpy()
    npp = 150
    pts = np.concatenate([_features(synth, npp, H, W, seed=30 + p) for p in range(3)])
    ga, gb = fe.make_pyramid(A, 4), fe.make_pyramid(B, 4)
    g = fe.track_fb(ga, gb, pts, pts, 4, n_per_pair=npp)
    for p in range(3):
        oa, ob = po.Pyramid(A[p], 4), po.Pyramid(B[p], 4)
        sl = slice(p * npp, (p + 1) * npp)
        o = po.hes_track_fb(oa, ob, pts[sl], pts[sl], 4)
        assert np.array_equal(g["accepted"][sl], o["accepted"])
        assert_bits_equal(g["to_xy"][sl], o["to_xy"], "pair %d to_xy" % p)
        assert_bits_equal(g["back_xy"][sl], o["back_xy"], "pair %d back_xy" % p)


def test_config2_measured_shape_bit_exact(fe, po, sfe, synth):
    """BASELINE config 2 exactly as bench.py runs it: 640x480 pairs, 2000 features per pair, a 4-level pyramid, several
    pairs per step in ONE pyramid batch of 2B slots (slots [0,B) = first frames, [B,2B) = second frames, built by one
    sfe_pyr_build_dev call), tracking addressed with from_first = 0 / to_first = B, device-resident inputs, and the
    2000 x 2000 Hamming match of each pair -- every output of every pair against a single-pair oracle run."""
    import importlib
    import torch
    bench = importlib.import_module("bench")
    B = 3
    W, H, NFEAT, LEVELS, DEPTH = bench.W, bench.H, bench.NFEAT, bench.LEVELS, bench.DEPTH
    assert (W, H, NFEAT, LEVELS, DEPTH) == (640, 480, 2000, 4, 4)
    dev = torch.device("cuda", 0)
    A, Bf, pts, q, t = bench.make_inputs(torch, synth, B, dev, seed=5)
    frames = torch.cat([A, Bf]).contiguous()
    n = B * NFEAT
    from_xy = torch.from_numpy(pts).to(dev)
    to_xy = from_xy.clone()
    fe.use_torch_stream()
    try:
        pyr = fe.pyramid(W, H, DEPTH, sfe.HESSIAN, 2 * B)
        pyr.build(frames)
        g = fe.track_fb(pyr, pyr, from_xy, to_xy, LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX, n_per_pair=NFEAT, from_first=0, to_first=B)
        q_d, t_d = torch.from_numpy(q.view(np.int32)).to(dev), torch.from_numpy(t.view(np.int32)).to(dev)
        idx, dist, ok = fe.match_hamming256(q_d, t_d, *bench.RATIO, batch=B)
        torch.cuda.synchronize()
    finally:
        fe.set_stream(None)
    g = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in g.items()}
    idx, dist, ok = (v.cpu().numpy() if hasattr(v, "cpu") else v for v in (idx, dist, ok))
    fr = frames.cpu().numpy()
    for p in range(B):
        oa, ob = po.Pyramid(fr[p], DEPTH), po.Pyramid(fr[B + p], DEPTH)
        for l in range(DEPTH):
            assert_bits_equal(pyr.plane(l, p), oa.plane(l), "pair %d first frame level %d" % (p, l))
            assert_bits_equal(pyr.plane(l, B + p), ob.plane(l), "pair %d second frame level %d" % (p, l))
        sl = slice(p * NFEAT, (p + 1) * NFEAT)
        o = po.hes_track_fb(oa, ob, pts[sl], pts[sl], LEVELS, bench.THR, bench.MAXIT, bench.FB_MAX)
        for k in ("status_fwd", "status_bwd", "accepted"):
            assert np.array_equal(g[k][sl], o[k]), (p, k)
        assert_bits_equal(g["to_xy"][sl], o["to_xy"], "pair %d to_xy" % p)
        assert_bits_equal(g["back_xy"][sl], o["back_xy"], "pair %d back_xy" % p)
        assert int(g["steps"][sl].sum()) == o["newton_steps"]
        oi, od, oo = po.hamming256_top2(q[sl], t[sl], *bench.RATIO)
        assert np.array_equal(idx[sl], oi) and np.array_equal(dist[sl], od) and np.array_equal(ok[sl], oo)
        assert o["accepted"].sum() > 0.5 * NFEAT


def test_replay_sequence_yuyv_equals_bgr_replay(fe, po, synth):
    """sfe_replay_sequence_yuyv: frames in the camera's native packed YUYV (2 bytes per pixel over PCIe), converted on
    the device with video.cpp:187-223's integer arithmetic, equal sfe_replay_sequence on the BGR frames that loop
    produces (oracle conversion), bit for bit."""
    H, W, S = 240, 320, 2
    rng = np.random.default_rng(9)
    A, B = synth.make_pairs(33, 4, H, W)
    bgr = np.concatenate([A.numpy(), B.numpy()])[[0, 1, 4, 5, 2, 3, 6, 7]]        # stride-2 pairs (0,4)->(0,2) etc.
    # a YUYV sequence whose luma follows the synthetic frames (chroma random): what matters is that both paths see
    # the BGR bytes video.cpp's loop makes of it
    yuyv = np.empty((len(bgr), H, W, 2), np.uint8)
    yuyv[..., 0] = bgr[..., 1]
    yuyv[..., 1] = rng.integers(96, 160, (len(bgr), H, W), dtype=np.uint8)
    conv = np.stack([po.yuyv_to_bgr(f.reshape(-1)).reshape(H, W, 3) for f in yuyv])
    npairs, npp = len(bgr) - S, 120
    pts = np.concatenate([synth.make_features(60 + p, npp, H, W, margin=16) for p in range(npairs)])
    a = fe.replay_sequence(np.ascontiguousarray(conv), S, pts, pts, 4, 3, n_per_pair=npp, chunk_pairs=4)
    b = fe.replay_sequence(yuyv, S, pts, pts, 4, 3, n_per_pair=npp, chunk_pairs=4)
    for k in ("status_fwd", "status_bwd", "accepted", "steps"):
        assert np.array_equal(a[k], b[k]), k
    assert_bits_equal(a["to_xy"], b["to_xy"], "to_xy")
    assert_bits_equal(a["back_xy"], b["back_xy"], "back_xy")
    o = po.hes_track_fb(po.Pyramid(conv[1], 4), po.Pyramid(conv[3], 4), pts[npp:2 * npp], pts[npp:2 * npp], 3)
    assert_bits_equal(b["to_xy"][npp:2 * npp], o["to_xy"], "pair 1 vs oracle")
    assert np.array_equal(b["accepted"][npp:2 * npp], o["accepted"])


@pytest.mark.parametrize("chunk", [2, 0, 5])
def test_replay_pairs_pipeline_matches_oracle(fe, po, synth, chunk):
    """sfe_replay_pairs (host buffers, chunk-pipelined: copy | pyramid (three sets) | tracking | download streams):
    9 pairs in chunks of 2 (five chunks, so every pyramid set and staging buffer is reused; ragged last chunk),
    automatic chunking and chunks of 5 -- every pair equals the single-pair oracle run bit for bit."""
    H, W = 240, 320
    npairs, npp = 9, 130
    A, B = synth.make_pairs(33, npairs, H, W)
    A, B = A.numpy(), B.numpy()
    pts = np.concatenate([_features(synth, npp, H, W, seed=40 + p) for p in range(npairs)])
    lv = np.where(np.arange(npairs * npp) % 4 == 0, 2, 4).astype(np.int32)
    g = fe.replay_pairs(A, B, pts, pts, depth=4, levels=lv, n_per_pair=npp, chunk_pairs=chunk)
    for p in range(npairs):
        oa, ob = po.Pyramid(A[p], 4), po.Pyramid(B[p], 4)
        sl = slice(p * npp, (p + 1) * npp)
        o = po.hes_track_fb(oa, ob, pts[sl], pts[sl], lv[sl])
        for k in ("status_fwd", "status_bwd", "accepted"):
            assert np.array_equal(g[k][sl], o[k]), (p, k)
        assert_bits_equal(g["to_xy"][sl], o["to_xy"], "pair %d to_xy" % p)
        assert_bits_equal(g["back_xy"][sl], o["back_xy"], "pair %d back_xy" % p)
    # pinned buffers + the asynchronous matcher queued ahead of the replay (the bench's e2e step)
    hq = synth.make_descriptors(5, 700, 0.2)
    ht = synth.make_descriptors(6, 900, 0.1)
    out = (np.empty((700, 2), np.int32), np.empty((700, 2), np.int32), np.empty(700, np.uint8))
    fe.match_hamming256_async(hq, ht, out, 4, 5, 80)
    pA, pB = fe.pinned(A.shape, np.uint8), fe.pinned(B.shape, np.uint8)
    pA[...] = A
    pB[...] = B
    g2 = fe.replay_pairs(pA, pB, pts, pts, depth=4, levels=lv, n_per_pair=npp, chunk_pairs=chunk)
    assert_bits_equal(g2["to_xy"], g["to_xy"], "pinned replay")
    fe.sync()  # the asynchronous matcher runs on a side stream of the context: its results are valid after sfe_sync()
    oi, od, oo = po.hamming256_top2(hq, ht, 4, 5, 80)
    assert np.array_equal(out[0], oi) and np.array_equal(out[1], od) and np.array_equal(out[2], oo)


@pytest.mark.parametrize("stride,chunk", [(2, 3), (1, 0), (2, 2), (3, 4)])
def test_replay_sequence_equals_replay_pairs(fe, synth, stride, chunk):
    """sfe_replay_sequence (one buffer of consecutive frames, pair i = (i, i + stride), every frame uploaded and built once
    per chunk) gives bit for bit what sfe_replay_pairs gives on the pairs spelled out -- which is checked against the
    oracle above -- incl. chunks shorter than the stride's halo and a ragged last chunk."""
    H, W, nframes, npp = 240, 320, 13, 90
    a, b = synth.make_pairs(71, nframes, H, W)
    frames = a.numpy().copy()
    frames[1::2] = b.numpy()[1::2]          # neighbouring frames related by the synthetic warp here and there
    npairs = nframes - stride
    pts = np.concatenate([_features(synth, npp, H, W, seed=90 + p) for p in range(npairs)])
    lv = np.where(np.arange(npairs * npp) % 3 == 0, 2, 4).astype(np.int32)
    ref = fe.replay_pairs(np.ascontiguousarray(frames[:npairs]), np.ascontiguousarray(frames[stride:]), pts, pts, depth=4,
                          levels=lv, n_per_pair=npp, chunk_pairs=chunk)
    got = fe.replay_sequence(frames, stride, pts, pts, depth=4, levels=lv, n_per_pair=npp, chunk_pairs=chunk)
    for k in ("status_fwd", "status_bwd", "accepted", "steps"):
        assert np.array_equal(got[k], ref[k]), k
    assert_bits_equal(got["to_xy"], ref["to_xy"], "to_xy")
    assert_bits_equal(got["back_xy"], ref["back_xy"], "back_xy")
    empty = fe.replay_sequence(frames[:stride], stride, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), depth=4, n_per_pair=npp)
    assert empty["to_xy"].shape == (0, 2)


def test_track_fb_empty_and_errors(fe, sfe, pair640):
    A, _ = pair640
    ga = fe.make_pyramid(A, 3)
    g = fe.track_fb(ga, ga, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), 3)
    assert g["to_xy"].shape == (0, 2)
    kp = fe.make_pyramid(A, 3, flavor=sfe.KLT)
    with pytest.raises(sfe.SlamFEError):
        fe.track_fb(kp, kp, np.ones((4, 2), np.float32), np.ones((4, 2), np.float32), 3)
    # identical frames: every interior feature must come back to where it started
    pts = np.float32([[100.25, 200.5], [320, 240], [600.75, 30.125]])
    g = fe.track_fb(ga, ga, pts, pts, 3)
    assert g["accepted"].all() and np.abs(g["to_xy"] - pts).max() < 1e-2


def test_klt_pyramid_and_track_bit_exact(fe, po, sfe, synth):
    H, W = 240, 320
    A, B = synth.make_pairs(4, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    n = 200
    pts = _features(synth, n, H, W, seed=9, border=0.1)
    ga, gb = fe.make_pyramid(A, 4, sfe.KLT), fe.make_pyramid(B, 4, sfe.KLT)
    oa, ob = po.Pyramid(A, 4, po.FLAVOR_KLT), po.Pyramid(B, 4, po.FLAVOR_KLT)
    g = fe.klt_track_fb(ga, gb, pts, pts)
    o = po.klt_track_fb(oa, ob, pts, pts)
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(g[k], o[k]), k
    assert_bits_equal(g["to_xy"], o["to_xy"], "klt to_xy")
    assert_bits_equal(g["back_xy"], o["back_xy"], "klt back_xy")
    assert int(g["steps"].sum()) == o["newton_steps"]
    # the symmetric-KLT normal equations (klt.h:286-353)
    xy = (pts + np.float32([0.4, -0.3])).astype(np.float32)
    for level in (0, 2):
        s = np.float32(0.5 ** level)
        sysg = fe.klt_system(ga, gb, level, pts * s, xy * s)
        for i in range(n):
            ref = po.klt_system(oa, pts[i, 0] * s, pts[i, 1] * s, ob, level, xy[i, 0] * s, xy[i, 1] * s)
            assert_bits_equal(sysg[i], ref, "klt system level %d #%d" % (level, i))


def test_brute_track_bit_exact(fe, po, sfe, synth):
    H, W = 240, 320
    A, B = synth.make_pairs(6, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    n = 48
    pts = synth.make_features(13, n, H, W, margin=20)
    pts[:4] = [[5, 100], [100, 8], [W - 6, 100], [100, H - 5]]  # inside the margin-13 band -> OUT_OF_BOUNDS
    ga, gb = fe.make_pyramid(A, 3, sfe.BRUTE), fe.make_pyramid(B, 3, sfe.BRUTE)
    oa, ob = po.Pyramid(A, 3, po.FLAVOR_BRUTE), po.Pyramid(B, 3, po.FLAVOR_BRUTE)
    g = fe.brute_track(ga, gb, pts, pts, fine=sfe.BRUTE_FINE_FAST)
    o = po.brute_track(oa, ob, pts, pts, fine=po.BRUTE_FINE_FAST)
    assert np.array_equal(g["status"], o["status"])
    assert_bits_equal(g["to_xy"], o["to_xy"], "brute to_xy")
    assert_bits_equal(g["best_sad"], o["best_sad"], "brute best_sad")
    assert g["positions"] == o["positions"]
    truth = synth.true_motion(pts, H, W, 0.004, (1.7, -2.3))
    ok = g["status"] == 0
    assert ok.sum() >= n - 8 and np.median(np.linalg.norm(g["to_xy"] - truth, axis=1)[ok]) < 0.15


def test_brute_track_reference_schedule_bit_exact(fe, po, sfe, synth):
    """BruteTracker::TrackFeature AS WRITTEN (brute.h:129-164): the default schedule includes the last level-0 pass
    SearchBest(stack[0], patches[0], 8, 0.01, &p) (brute.h:158), 1600 x 1600 positions per feature, which decides the
    returned position and the sad of the `sad > 100` gate.  A handful of features, bit-exact against the oracle, which
    is itself pinned to a tier-0 (cv2.getRectSubPix) run of the same schedule by tests/golden/brute_tracks.npz."""
    H, W = 97, 131
    A, B = synth.make_pairs(42, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    pts = synth.make_features(12, 6, H, W, margin=20)
    pts[5] = [W - 14.5, 30.25]  # the 8 px window reaches past the right border: the tile holds replicated columns
    assert len(sfe.BRUTE_FINE) == 10 and tuple(sfe.BRUTE_FINE[-2:]) == (8.0, np.float32(0.01))
    ga, gb = fe.make_pyramid(A, 3, sfe.BRUTE), fe.make_pyramid(B, 3, sfe.BRUTE)
    oa, ob = po.Pyramid(A, 3, po.FLAVOR_BRUTE), po.Pyramid(B, 3, po.FLAVOR_BRUTE)
    g = fe.brute_track(ga, gb, pts, pts)          # defaults = the reference's schedule
    o = po.brute_track(oa, ob, pts, pts)
    assert np.array_equal(g["status"], o["status"])
    assert_bits_equal(g["to_xy"], o["to_xy"], "brute to_xy (reference schedule)")
    assert_bits_equal(g["best_sad"], o["best_sad"], "brute best_sad (reference schedule)")
    assert g["positions"] == o["positions"] and g["positions"] >= 6 * 1600 * 1600


@pytest.fixture(params=["alu", "mma"])
def hamming_kernel(request, fe):
    """Runs a test once per matcher kernel: the integer-ALU kernel (hamming.cu) and the tcgen05 int8 contraction
    (hamming_mma.cu) must both be bit-identical to the oracle."""
    prev = fe.hamming_impl({"alu": 1, "mma": 2}[request.param])
    yield request.param
    fe.hamming_impl(prev)


@pytest.mark.parametrize("nq,nt,batch", [(2000, 2000, 1), (777, 1301, 3), (1, 1, 1), (5, 0, 1), (300, 1, 2), (4096, 70000, 1),
                                         (128, 256, 1), (129, 257, 2), (500, 513, 40), (3000, 255, 1)])
def test_hamming_bit_exact(fe, po, synth, hamming_kernel, nq, nt, batch):
    t = synth.make_descriptors(1, max(nt * batch, 1), dup_frac=0.05)[:nt * batch]
    q = synth.make_descriptors(2, nq * batch, dup_frac=0.3, source=t if nt else None)
    idx, dist, ok = fe.match_hamming256(q, t, 4, 5, 80, batch=batch)
    for b in range(batch):
        oi, od, oo = po.hamming256_top2(q[b * nq:(b + 1) * nq], t[b * nt:(b + 1) * nt], 4, 5, 80)
        sl = slice(b * nq, (b + 1) * nq)
        assert np.array_equal(idx[sl], oi) and np.array_equal(dist[sl], od) and np.array_equal(ok[sl], oo)


def test_config3_1080p_8_levels_5000_features_bit_exact(fe, po, synth):
    """BASELINE config 3: 1920x1080 pair, 5000 features, 8-level pyramid (1920 ... 15x9) -- pyramids of both
    frames and the forward/backward tracks against the oracle, bit for bit."""
    H, W, n = 1080, 1920, 5000
    A, B = synth.make_pairs(5, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    pts = _features(synth, n, H, W, seed=21, border=0.05)
    ga, gb = fe.make_pyramid(A, 8), fe.make_pyramid(B, 8)
    oa, ob = po.Pyramid(A, 8), po.Pyramid(B, 8)
    for l in range(8):
        assert_bits_equal(gb.plane(l), ob.plane(l), "frame B level %d" % l)
    g = fe.track_fb(ga, gb, pts, pts, 8)
    o = po.hes_track_fb(oa, ob, pts, pts, 8)
    for k in ("status_fwd", "status_bwd", "accepted"):
        assert np.array_equal(g[k], o[k]), k
    assert_bits_equal(g["to_xy"], o["to_xy"], "to_xy")
    assert_bits_equal(g["back_xy"], o["back_xy"], "back_xy")
    assert int(g["steps"].sum()) == o["newton_steps"]
    assert g["accepted"].mean() > 0.8


def test_hamming_extreme_rows_bit_exact(fe, po, synth, hamming_kernel):
    """All-zero / all-one rows, distances 0 and 256, every tie broken by the lowest train index: the corners of the
    accumulator range of the int8 contraction (8192 d + j - 128 - 2^20, d = 0 ... 256, j = 0 ... 255)."""
    rng = np.random.default_rng(3)
    t = synth.make_descriptors(5, 1024, dup_frac=0.0)
    t[0] = 0; t[1] = 0xffffffff; t[255] = 0; t[256] = 0xffffffff; t[511] = t[700]; t[1023] = 0
    q = synth.make_descriptors(6, 640, dup_frac=0.0)
    q[0] = 0; q[1] = 0xffffffff; q[2] = t[700]; q[3] = ~t[700]; q[639] = 0xffffffff
    q[4:132] = t[rng.integers(0, 1024, 128)]
    idx, dist, ok = fe.match_hamming256(q, t, 4, 5, 80)
    oi, od, oo = po.hamming256_top2(q, t, 4, 5, 80)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od) and np.array_equal(ok, oo)
    assert tuple(idx[0]) == (0, 255) and tuple(dist[0]) == (0, 0) and tuple(idx[2]) == (511, 700)
    # one train tile whose every row ties at the maximum distance
    t2 = np.zeros((300, 8), np.uint32)
    q2 = np.full((130, 8), 0xffffffff, np.uint32)
    i2, d2, _ = fe.match_hamming256(q2, t2, 4, 5, 80)
    assert (i2 == np.array([0, 1])).all() and (d2 == 256).all()


def test_hamming_random_shapes_bit_exact(fe, po, synth, hamming_kernel):
    """Randomly drawn problem shapes: partial query tiles, partial train tiles, one to many train splits per query tile
    (few queries against a long train set puts every SM on its own slice of the train rows), batches -- each against
    the oracle, for both kernels."""
    rng = np.random.default_rng(2026)
    shapes = [(int(rng.integers(1, 700)), int(rng.integers(1, 3000)), int(rng.integers(1, 5))) for _ in range(10)]
    shapes += [(100, 100000, 1), (3, 40000, 2), (1000, 7, 3), (257, 511, 1)]
    for nq, nt, batch in shapes:
        t = synth.make_descriptors(nq + 7 * nt, nt * batch, dup_frac=0.05)
        q = synth.make_descriptors(nq * 3 + nt, nq * batch, dup_frac=0.3, source=t)
        idx, dist, ok = fe.match_hamming256(q, t, 4, 5, 80, batch=batch)
        for b in range(batch):
            oi, od, oo = po.hamming256_top2(q[b * nq:(b + 1) * nq], t[b * nt:(b + 1) * nt], 4, 5, 80)
            sl = slice(b * nq, (b + 1) * nq)
            assert np.array_equal(idx[sl], oi) and np.array_equal(dist[sl], od) and np.array_equal(ok[sl], oo), (nq, nt, batch, b)


def test_config5_hamming_1m_x_1m_properties(fe, po, synth, hamming_kernel):
    """BASELINE config 5 at full size on one GPU (10^12 comparisons): the oracle cannot finish that, so the result is
    checked through size-independent properties -- planted duplicates are found at distance 0 at the LOWEST train index
    that holds them, distances are ordered, the ratio flag is consistent with the returned distances -- and 256 randomly
    chosen query rows are compared with the oracle outright."""
    n = 1 << 20
    t = synth.make_descriptors(11, n, dup_frac=0.001)
    q = synth.make_descriptors(12, n, dup_frac=0.0)
    rng = np.random.default_rng(7)
    planted = rng.choice(n, 4096, replace=False)
    src = rng.integers(0, n, 4096)
    q[planted] = t[src]
    idx, dist, ok = fe.match_hamming256(q, t, 4, 5, 80)
    assert idx.shape == (n, 2) and (idx >= 0).all() and (idx < n).all()
    assert (dist[:, 0] <= dist[:, 1]).all() and (dist >= 0).all() and (dist <= 256).all()
    assert (idx[:, 0] != idx[:, 1]).all()
    # planted rows: distance 0, and the winner holds the same descriptor at an index <= the planted source
    assert (dist[planted, 0] == 0).all()
    assert (idx[planted, 0] <= src).all()
    assert (t[idx[planted, 0]] == t[src]).all()
    # distances recomputed on the host for every query row agree with the returned indices
    pop = np.array([bin(i).count("1") for i in range(256)], np.uint8)
    for col in (0, 1):
        x = (q ^ t[idx[:, col]]).view(np.uint8)
        assert np.array_equal(pop[x].reshape(n, 32).sum(1).astype(np.int32), dist[:, col]), "column %d" % col
    exp_ok = (dist[:, 0] <= 80) & (dist[:, 0].astype(np.int64) * 5 < dist[:, 1].astype(np.int64) * 4)
    assert np.array_equal(ok.astype(bool), exp_ok)
    # spot check against the oracle
    rows = np.concatenate([planted[:64], rng.choice(n, 192, replace=False)])
    oi, od, oo = po.hamming256_top2(q[rows], t, 4, 5, 80)
    assert np.array_equal(idx[rows], oi) and np.array_equal(dist[rows], od) and np.array_equal(ok[rows], oo)


@pytest.mark.parametrize("shape", [(480, 640), (240, 320), (96, 83), (33, 48), (1080, 1920)])
def test_good_features_bit_exact(fe, po, synth, shape):
    """goodFeaturesToTrack (matcher.cpp:123-130) on the RGB2GRAY image: response map bit for bit, corner lists
    identical (order included) for the reference's parameters and for denser settings."""
    H, W = shape
    frames = synth.make_frames(17 + H, 2, H, W).numpy()
    for (maxc, q, mind) in ((120, 0.01, 20.0), (2000, 0.01, 5.0), (500, 0.001, 3.5), (300, 0.05, 0.0)):
        corners, eig = fe.good_features(frames, maxc, q, mind, want_eig=True)
        for f in range(2):
            oc, oe, _ = po.good_features(frames[f], maxc, q, mind, want_eig=True)
            assert_bits_equal(eig[f], oe, "response map %dx%d frame %d" % (W, H, f))
            assert corners[f].shape == oc.shape and np.array_equal(corners[f], oc), (shape, maxc, q, mind, f, len(corners[f]), len(oc))


def test_good_features_device_batch(fe, po, synth):
    """Device-pointer variant on a batch; a flat (textureless) frame yields no corners."""
    import torch
    H, W = 240, 320
    frames = synth.make_frames(3, 6, H, W)
    frames[4] = 77
    xy, cnt = fe.good_features(frames.cuda(), 120, 0.01, 20.0)
    fe.sync()
    xy, cnt = xy.cpu().numpy(), cnt.cpu().numpy()
    for f in range(6):
        oc = po.good_features(frames[f].numpy(), 120, 0.01, 20.0)
        assert cnt[f] == len(oc) and np.array_equal(xy[f, :cnt[f]], oc), f
    assert cnt[4] == 0


def test_seed_features_bit_exact(fe, po):
    """Head of the FindMatches loop (matcher.cpp:224-245): levels from the uncertainty, Frame::Project seeds
    (project.h:11-54, doubles) and the out-of-bounds gate -- incl. points behind the lens and on the bounds."""
    rng = np.random.default_rng(9)
    n = 5000
    ang = 0.3
    axis = np.float64([0.2, -0.5, 0.1]); axis /= np.linalg.norm(axis)
    rot = np.concatenate([axis * np.sin(ang / 2), [np.cos(ang / 2)]])
    trans = np.float64([0.3, -0.1, 0.5])
    k = np.float64([-0.12, 0.03, -0.004, 310.0, 305.0, 322.5, 238.25])
    pts = np.concatenate([rng.normal(0, 3, (n, 2)), rng.uniform(-2, 12, (n, 1)), rng.uniform(0.5, 2, (n, 1))], 1)
    unc = rng.choice([1e8, 150.0, 100.0, 99.0, 3.0], n)
    from_xy = np.stack([rng.uniform(-5, 650, n), rng.uniform(-5, 490, n)], 1).astype(np.float32)
    from_xy[:4] = [[0, 0], [640, 100], [100, 480], [639.99, 479.99]]
    unc[:4] = 1e8
    gs, gl, gg = fe.seed_features(pts, unc, rot, trans, k, from_xy, 640, 480)
    os_, ol, og = po.seed_features(pts, unc, rot, trans, k, from_xy, 640, 480)
    assert_bits_equal(gs, os_, "seeds")
    assert np.array_equal(gl, ol) and np.array_equal(gg, og)
    assert set(np.unique(gl)) == {3, 6} and 0 < gg.mean() < 1
    assert list(gg[:4]) == [1, 0, 1, 1]   # x >= cols is out, y == rows is still in (matcher.cpp:243 uses `>` on y)


def test_yuyv_to_bgr_bit_exact(fe, po):
    """video.cpp:187-223 on a random VGA buffer plus the saturating corners of the colour cube."""
    rng = np.random.default_rng(2)
    buf = rng.integers(0, 256, 640 * 480 * 2, dtype=np.uint8)
    buf[:16] = [0, 0, 0, 0, 255, 255, 255, 255, 255, 0, 255, 0, 0, 255, 0, 255]
    assert np.array_equal(fe.yuyv_to_bgr(buf), po.yuyv_to_bgr(buf))
    assert fe.yuyv_to_bgr(np.zeros(0, np.uint8)).size == 0


@pytest.mark.parametrize("shape,pad", [((96, 640), 4), ((60, 1920), 4), ((48, 1280), 12), ((40, 640), 1), ((64, 1920), 16)])
def test_pyramid_device_frames_with_row_padding(fe, po, synth, shape, pad):
    """sfe_pyr_build_dev on caller-owned device frames whose rows are padded: a stride that is a multiple of 4 but not of
    16 takes the row kernel's per-lane cp.async ring instead of the bulk copies (640 / 1280 columns, and the four column
    segments of 1920), a multiple of 16 the bulk copies with a non-dense stride, an odd stride the tiled kernels."""
    import torch
    H, W = shape
    frames = synth.make_frames(31, 2, H, W)
    stride = W * 3 + pad
    buf = torch.zeros((2, H, stride), dtype=torch.uint8, device="cuda")
    buf[:, :, :W * 3] = frames.reshape(2, H, W * 3).cuda()
    torch.cuda.synchronize()
    gp = fe.pyramid(W, H, 3, 0, 2)
    assert fe.L.sfe_pyr_build_dev(fe.h, gp.h, buf.data_ptr(), stride, stride * H, 0, 2) == 0
    fe.sync()
    for f in range(2):
        op = po.Pyramid(frames[f].numpy(), 3, 0)
        for l in range(3):
            assert_bits_equal(gp.plane(l, f), op.plane(l), "%dx%d pad %d frame %d level %d" % (W, H, pad, f, l))
