#!/usr/bin/env python
"""Generates tests/golden/*.npz with the REAL OpenCV (cv2) through oracle/tier0_cv2.py.

Run in the build container (cv2 4.13 is installed there):   python tests/golden/make_golden.py
The fixtures pin the dependency-free C oracle (oracle/oracle.c) -- and through it the CUDA path --
against outputs of the OpenCV primitives the reference calls (hessian.h / klt.h / brute.h):

  pyramid_*.npz   cv2-built pyramids (all three flavours) of a seeded 131x97 BGR frame
  patches.npz     cv2.getRectSubPix-based GetPatch outputs (hessian.h:54-93) at interior, edge and
                  corner centres, incl. the top-right quirk region
  hessian.npz     BruteHessian 7-tuples (hessian.h:147-172) from tier-0 on the golden planes
  tracks.npz      forward/backward tracks (matcher.cpp:173-206) from tier-0 ON THE cv2 PLANES
                  (the oracle must reproduce them bit-exactly when given the same planes)
  hamming.npz     cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) on seeded descriptors with planted ties
  corners.npz     cv2.cornerMinEigenVal / cv2.goodFeaturesToTrack on the RGB2GRAY image of seeded frames
                  (matcher.cpp:313 + :123-130), for the reference's parameters (120, 0.01, 20) and denser ones
  pyramid640.npz  SHA-256 of every plane of the cv2-built 6-level Hessian, 3-level KLT and 3-level brute pyramids of a
                  seeded 640x480 frame and of a 1920x1080 frame's 8-level Hessian pyramid (the measured shapes; the
                  planes themselves are 1.6 / 11 MB, a digest pins every bit), plus the tail columns in full
  klt_tracks.npz  klt.h on cv2 planes: per Newton iteration the point and the 24 numbers
                  A,B,C,RS,VW,U,e,d of klt.h:286-343, the six finite differences, and the forward/backward results
  brute_tracks.npz  brute.h on cv2 planes: the arg-min and score after every SearchBest pass,
                  for the four cheap passes on 8 features and for the schedule AS WRITTEN (incl. (8, 0.01)) on 2

It also re-runs the arithmetic probes that fixed the oracle's operation order (see oracle.h) and
prints what fraction of each cv2 primitive's output the oracle reproduces bit-for-bit.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import pyoracle as po  # noqa: E402
from oracle import tier0_cv2 as t0  # noqa: E402

synth = importlib.import_module("slam-robot_b200.synth")


def probes():
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    g = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY)
    print("cvtColor RGB2GRAY   bit-exact:", np.array_equal(g, po.gray_u8(bgr)))
    codes = np.arange(256, dtype=np.uint8).reshape(1, 256)
    conv = t0.convert_to_via_normalize(codes)
    print("convertTo(1/255.)   bit-exact:", np.array_equal(conv, codes.astype(np.float32) * np.float32(1 / 255.)))
    for sigma in (1.1, 0.8, 0.6):
        k = cv2.getGaussianKernel(5, sigma, cv2.CV_32F).ravel()
        print("gaussian taps sigma=%.1f: %s" % (sigma, [hex(int(x.view(np.uint32))) for x in k[:3]]))
    for (h, w) in ((96, 128), (97, 131), (15, 20), (17, 30), (9, 15), (33, 66), (25, 33)):
        img = rng.random((h, w), dtype=np.float32)
        for sigma in (1.1, 0.8, 0.6):
            a, b = po.gauss5(img, sigma), cv2.GaussianBlur(img, (5, 5), sigma, sigmaY=sigma)
            print("GaussianBlur %.1f %dx%d: bit-exact incl. scalar-tail columns: %s" % (
                sigma, w, h, np.array_equal(a.view(np.uint32), b.view(np.uint32))))
        a, b = po.pyrdown(img), cv2.pyrDown(img)
        print("pyrDown %dx%d: bit-exact incl. border/scalar columns: %s" % (w, h, np.array_equal(a.view(np.uint32), b.view(np.uint32))))
        gx, gy = po.scharr(img)
        cx = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=cv2.FILTER_SCHARR, scale=1 / 32.)
        cy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=cv2.FILTER_SCHARR, scale=1 / 32.)
        print("Scharr %dx%d: bit-exact: %s %s" % (w, h, np.array_equal(gx.view(np.uint32), cx.view(np.uint32)),
                                                  np.array_equal(gy.view(np.uint32), cy.view(np.uint32))))
    img = rng.random((48, 64), dtype=np.float32) + np.float32(0.1)
    bad = 0
    for _ in range(3000):
        n, m = int(rng.integers(6, 14)), int(rng.integers(6, 14))
        cx, cy = float(np.float32(rng.uniform(-3, 67))), float(np.float32(rng.uniform(-3, 51)))
        bad += not np.array_equal(cv2.getRectSubPix(img, (n, m), (cx, cy)), po.rect_subpix(img, n, m, cx, cy))
    print("getRectSubPix: %d of 3000 random windows (incl. all borders/corners) differ" % bad)


def digest(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8).copy()


def measured_shapes():
    """cv2 pyramids at the sizes bench.py measures (C2/C3); frames are regenerated from their seeds by the tests."""
    out = {}
    for name, (h, w, seed, depth) in {"vga": (480, 640, 1, 6), "hd": (1080, 1920, 2, 8)}.items():
        fr = synth.make_frames(seed, 1, h, w).numpy()[0]
        out[name + "_seed"] = np.int64([seed, h, w, depth])
        out[name + "_gray_sum"] = np.int64(cv2.cvtColor(fr, cv2.COLOR_RGB2GRAY).astype(np.int64).sum())
        hes = t0.pyramid_hessian(fr, depth)
        for l, p in enumerate(hes):
            out["%s_hes%d_sha" % (name, l)] = digest(p)
            out["%s_hes%d_tail" % (name, l)] = np.concatenate([p[:, :1], p[:, -8:]], axis=1)  # col 0 + last 8 columns
        if name == "vga":
            for l, planes in enumerate(t0.pyramid_klt(fr, 3)):
                for k in range(3):
                    out["vga_klt%d_%d_sha" % (l, k)] = digest(planes[k])
            for l, p in enumerate(t0.pyramid_brute(fr, 3)):
                out["vga_bru%d_sha" % l] = digest(p)
        o = po.Pyramid(fr, depth, po.FLAVOR_HESSIAN)
        print("%dx%d Hessian pyramid, %d levels: oracle == cv2 bit for bit: %s" % (
            w, h, depth, all(np.array_equal(o.plane(l).view(np.uint32), hes[l].view(np.uint32)) for l in range(depth))))
    np.savez_compressed(os.path.join(HERE, "pyramid640.npz"), **out)


def klt_and_brute():
    H, W = 97, 131
    A, B = synth.make_pairs(42, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    # (v) klt.h: per-iteration systems on the cv2 planes
    ka, kb = t0.pyramid_klt(A, 3), t0.pyramid_klt(B, 3)
    pts = synth.make_features(11, 12, H, W, margin=10, border_frac=0.25)
    traces = []
    r = t0.klt_track_fb(ka, kb, pts, traces=traces)
    feat, lvl, xy, sys24, d6 = [], [], [], [], []
    for i, tr in enumerate(traces):
        for (level, its) in tr:
            for (x, y, s24, dd) in its:
                feat.append(i); lvl.append(level); xy.append((x, y)); sys24.append(s24); d6.append(dd)
    np.savez_compressed(os.path.join(HERE, "klt_tracks.npz"), xy=pts, it_feature=np.int32(feat), it_level=np.int32(lvl),
                        it_xy=np.float32(xy), it_sys24=np.float32(sys24), it_d6=np.float32(d6),
                        **{k: np.asarray(v) for k, v in r.items()})
    print("klt tier-0: %d Newton iterations recorded, %d/12 accepted" % (len(feat), int(r["accepted"].sum())))
    # (vi) brute.h: SearchBest arg-mins on the cv2 planes
    ba, bb = t0.pyramid_brute(A, 3), t0.pyramid_brute(B, 3)
    pts = synth.make_features(12, 8, H, W, margin=20)
    pts[5] = [W - 14.5, 30.25]
    out = dict(xy=pts)
    for tag, fine, count in (("fast", t0.BRUTE_FINE[:4], 8), ("ref", t0.BRUTE_FINE, 2)):
        st, fin, passes = [], [], []
        for i in range(count):
            tr = []
            s, x, y, sad = t0.brute_track_feature(ba, pts[i, 0], pts[i, 1], bb, pts[i, 0], pts[i, 1], fine=fine, trace=tr)
            st.append(s); fin.append((x, y, sad)); passes.append([t[3:6] for t in tr])
            npos = sum(t[6] for t in tr)
        out[tag + "_status"], out[tag + "_final"], out[tag + "_passes"] = np.int32(st), np.float32(fin), np.float32(passes)
        out[tag + "_positions_per_feature"] = np.int64(npos)
        print("brute tier-0 (%s schedule): %d features, %d positions each" % (tag, count, npos))
    np.savez_compressed(os.path.join(HERE, "brute_tracks.npz"), **out)


def main():
    probes()
    H, W = 97, 131
    A, B = synth.make_pairs(42, 1, H, W)
    A, B = A[0].numpy(), B[0].numpy()
    # pyramids from real OpenCV
    hes_a, hes_b = t0.pyramid_hessian(A, 4), t0.pyramid_hessian(B, 4)
    klt_a = t0.pyramid_klt(A, 3)
    bru_a = t0.pyramid_brute(A, 3)
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), A=A, B=B,
                        **{"hes_a%d" % i: p for i, p in enumerate(hes_a)}, **{"hes_b%d" % i: p for i, p in enumerate(hes_b)},
                        **{"klt_a%d_%d" % (i, k): p[k] for i, p in enumerate(klt_a) for k in range(3)},
                        **{"bru_a%d" % i: p for i, p in enumerate(bru_a)})
    # patches (hessian.h GetPatch through cv2.getRectSubPix)
    rng = np.random.default_rng(7)
    cent, lev, pat, mean, sumsq = [], [], [], [], []
    for level in range(4):
        h, w = hes_a[level].shape
        pts = np.stack([rng.uniform(0.02, w - 0.02, 60), rng.uniform(0.02, h - 0.02, 60)], 1).astype(np.float32)
        pts[:8] = [[0.5, 0.5], [w - 0.5, 0.5], [0.5, h - 0.5], [w - 0.5, h - 0.5], [w - 3.3, 0.2], [w - 6.9, 5.7], [6.4, 6.4], [6.6, 6.6]]
        for (x, y) in pts:
            d, m, q = t0.hes_get_patch(hes_a[level], x, y)
            cent.append((x, y)); lev.append(level); pat.append(d); mean.append(m); sumsq.append(q)
    np.savez_compressed(os.path.join(HERE, "patches.npz"), xy=np.float32(cent), level=np.int32(lev), patch=np.float32(pat),
                        mean=np.float32(mean), sumsq=np.float32(sumsq))
    # BruteHessian tuples
    pts = synth.make_features(3, 40, H, W, margin=10, border_frac=0.3)
    tup = []
    for (x, y) in pts:
        d, m, q = t0.hes_get_patch(hes_a[0], x, y)
        s0, d6 = t0.hes_brute_hessian(hes_b[0], d, m, q, np.float32(x + 0.7), np.float32(y - 0.4))
        tup.append([s0] + list(d6))
    np.savez_compressed(os.path.join(HERE, "hessian.npz"), xy=pts, out7=np.float32(tup))
    # end-to-end forward/backward tracks on the cv2 planes
    pts = synth.make_features(5, 80, H, W, margin=8, border_frac=0.25)
    r3 = t0.hes_track_fb(hes_a, hes_b, pts, pts, 3)
    r4 = t0.hes_track_fb(hes_a, hes_b, pts, pts, 4)
    np.savez_compressed(os.path.join(HERE, "tracks.npz"), xy=pts,
                        **{"l3_" + k: np.asarray(v) for k, v in r3.items()}, **{"l4_" + k: np.asarray(v) for k, v in r4.items()})
    print("tracks: accepted %d/80 (3 levels), %d/80 (4 levels)" % (r3["accepted"].sum(), r4["accepted"].sum()))
    # Hamming knn with planted ties
    t = synth.make_descriptors(1, 700, dup_frac=0.1)
    q = synth.make_descriptors(2, 500, dup_frac=0.4, source=t)
    idx, dist = t0.hamming_knn2_cv2(q.view(np.uint8), t.view(np.uint8))
    print("hamming: %d queries with tied best distance" % int((dist[:, 0] == dist[:, 1]).sum()))
    np.savez_compressed(os.path.join(HERE, "hamming.npz"), q=q, t=t, idx=idx, dist=dist)
    # corner seeding: response map and corner lists from the real OpenCV
    out = {}
    params = [(120, 0.01, 20.0), (2000, 0.01, 5.0), (500, 0.001, 3.5), (300, 0.05, 0.0)]
    out["params"] = np.float64(params)
    for name, (h, w, seed) in {"s": (96, 128, 3), "m": (240, 320, 4), "vga": (480, 640, 1)}.items():
        fr = synth.make_frames(seed, 1, h, w).numpy()[0]
        grey = cv2.cvtColor(fr, cv2.COLOR_RGB2GRAY)
        eig = cv2.cornerMinEigenVal(grey, 3, ksize=3)
        if name != "vga":
            out[name + "_frame"] = fr
            out[name + "_eig"] = eig
        else:  # the VGA frame is regenerated from its seed; its gray image is stored as a checksum
            out[name + "_seed"] = np.int64([seed, h, w])
            out[name + "_gray_sum"] = np.int64(grey.astype(np.int64).sum())
            out[name + "_eig_rows"] = eig[::40]
        for k, (maxc, q, mind) in enumerate(params):
            c = cv2.goodFeaturesToTrack(grey, maxc, q, mind)
            out["%s_corners%d" % (name, k)] = c.reshape(-1, 2) if c is not None else np.zeros((0, 2), np.float32)
        oc, oe, _ = po.good_features(fr, 120, 0.01, 20.0, want_eig=True)
        print("goodFeaturesToTrack %dx%d: response map bit-exact: %s, corners identical: %s" % (
            w, h, np.array_equal(oe.view(np.uint32), eig.view(np.uint32)), np.array_equal(oc, out[name + "_corners0"])))
    np.savez_compressed(os.path.join(HERE, "corners.npz"), **out)
    measured_shapes()
    klt_and_brute()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-14s %7.1f KB" % (f, os.path.getsize(os.path.join(HERE, f)) / 1024))


if __name__ == "__main__":
    main()
