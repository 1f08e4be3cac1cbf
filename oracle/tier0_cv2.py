"""Tier-0 oracle: the reference's front-end restated in Python ON TOP OF THE REAL OpenCV.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Every OpenCV primitive the reference calls
(cvtColor, convertTo, GaussianBlur, pyrDown, Sobel/Scharr, getRectSubPix) is executed by cv2
itself; only the reference's own control flow and arithmetic (hessian.h, klt.h, brute.h,
matcher.cpp:173-206) is restated, in NumPy float32 with the declared lane/tree summation
order, and double where the reference uses double.

Used for two things: (1) pinning the dependency-free C restatement (oracle.c) against real
OpenCV outputs, (2) generating the golden fixtures under tests/golden/ (make_golden.py).
cv2 is optional at test time: tests that need it skip when it cannot be imported.
"""
import math

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

f32 = np.float32
OK, SMALL_DET, OUT_OF_BOUNDS = 0, 1, 2
N = 13
LEN = 169


def have_cv2():
    return cv2 is not None


def fma(a, b, c):
    """float32 fused multiply-add, emulated through float64 (a*b is exact in f64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def tree32(lanes):
    p = np.array(lanes, dtype=np.float32)
    off = 16
    while off:
        p[:off] = p[:off] + p[off:2 * off]
        off >>= 1
    return f32(p[0])


def tree32_adj(lanes):
    """The Hessian tracker's tree (oracle.h): adjacent lanes first, strides 1,2,4,8,16."""
    p = np.array(lanes, dtype=np.float32)
    off = 1
    while off < 32:
        p[::2 * off] = p[::2 * off] + p[off::2 * off]
        off <<= 1
    return f32(p[0])


def lane_sum(vals, acc_fn, adjacent=False):
    """Declared reduction: pixel i -> lane i%32, sequential within a lane, then tree32 (klt.h, brute.h) or
    tree32_adj (hessian.h)."""
    vals = np.asarray(vals, np.float32)
    lanes = np.zeros(32, np.float32)
    for k in range(0, LEN, 32):
        chunk = vals[k:k + 32]
        lanes[:len(chunk)] = acc_fn(lanes[:len(chunk)], chunk, k)
    return tree32_adj(lanes) if adjacent else tree32(lanes)


def mask13():
    """hessian.h:11-30"""
    m = np.empty(LEN, np.float32)
    for y in range(N):
        for x in range(N):
            rx = 0.5 * N - x
            ry = 0.5 * N - y
            m[y * N + x] = 1. / (15. + rx * rx + ry * ry)
    s = 0.0
    for v in m:
        s += float(v)
    scale = LEN / s
    return (m.astype(np.float64) * scale).astype(np.float32)


_MASK = None


def _mask():
    global _MASK
    if _MASK is None:
        _MASK = mask13()
    return _MASK


# ------------------------------------------------------------------ pyramids (all cv2)

def gray_f32(bgr):
    g = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY)  # hessian.h:100 (RGB2GRAY applied to BGR bytes)
    # Mat::convertTo(CV_32F, 1/255.) has no Python binding; cv2.normalize(MINMAX 0..1) calls it with
    # exactly that scale when the image spans 0..255, and equals this float32 product on all codes.
    return g.astype(np.float32) * f32(1. / 255.)


def convert_to_via_normalize(codes_u8):
    """Real Mat::convertTo(CV_32F, 1/255.) reached through cv2.normalize (input must span 0..255)."""
    return cv2.normalize(codes_u8, None, 1, 0, cv2.NORM_MINMAX, dtype=cv2.CV_32F)


def pyramid_hessian(bgr, depth):
    """hessian.h:95-126"""
    img = cv2.GaussianBlur(gray_f32(bgr), (5, 5), 1.1, sigmaY=1.1)
    out = [img]
    for _ in range(1, depth):
        out.append(cv2.GaussianBlur(cv2.pyrDown(out[-1]), (5, 5), .8, sigmaY=.8))
    return out


def pyramid_klt(bgr, depth):
    """klt.h:98-137 -> list of (img, gx, gy)"""
    img = gray_f32(bgr)
    gx = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=cv2.FILTER_SCHARR, scale=1. / 32.)
    gy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=cv2.FILTER_SCHARR, scale=1. / 32.)
    out = [(img, gx, gy)]
    for _ in range(1, depth):
        pi, pgx, pgy = out[-1]
        ci = cv2.GaussianBlur(cv2.pyrDown(pi), (5, 5), .6, sigmaY=.6)
        out.append((ci, cv2.pyrDown(pgx) * f32(2), cv2.pyrDown(pgy) * f32(2)))
    return out


def pyramid_brute(bgr, depth):
    """brute.h:59-80"""
    out = [gray_f32(bgr)]
    for _ in range(1, depth):
        out.append(cv2.pyrDown(out[-1]))
    return out


# ------------------------------------------------------------------ patches

def patch_stats(data, adjacent=False):
    d = np.asarray(data, np.float32).ravel()
    s = lane_sum(d, lambda acc, x, k: acc + x, adjacent)
    q = lane_sum(d, lambda acc, x, k: fma(x, x, acc), adjacent)
    return f32(s / f32(LEN)), f32(q / f32(LEN))


def hes_get_patch(img, x, y):
    """hessian.h:54-93 with cv2.getRectSubPix"""
    px, py = f32(x), f32(y)
    data = np.zeros((N, N), np.float32)
    rx = ry = 0
    rw = rh = N
    if float(px) < 0.5 * N:
        d = int((0.5 * N - float(px)) + 0.9999)
        px = f32(float(px) + 0.5 * d)
        rx, rw = d, N - d
    if float(py) < 0.5 * N:
        d = int(0.5 * N - float(py))
        py = f32(float(py) + 0.5 * d)
        ry, rh = d, N - d
    if rw > 0 and rh > 0:
        data[ry:ry + rh, rx:rx + rw] = cv2.getRectSubPix(img, (rw, rh), (float(px), float(py)))
    m, q = patch_stats(data, adjacent=True)
    return data, m, q


def full_patch(img, x, y):
    """klt.h:72-76 / brute.h:43-47"""
    data = cv2.getRectSubPix(img, (N, N), (float(f32(x)), float(f32(y))))
    m, q = patch_stats(data)
    return data, m, q


def hes_score(p1, m1, q1, p2, m2, q2):
    """hessian.h:129-141"""
    p1 = np.asarray(p1, np.float32).ravel()
    p2 = np.asarray(p2, np.float32).ravel()
    with np.errstate(all="ignore"):
        alpha = f32(np.sqrt(f32(q1) / f32(q2)))
        beta = f32(f32(m1) - alpha * f32(m2))
        diff = fma(-p2, alpha, p1) - beta
        term = diff * diff
    keep = (p1 != 0) & (p2 != 0)
    mask = _mask()

    def acc_fn(acc, x, k):
        kk = keep[k:k + len(x)]
        return np.where(kk, fma(x, mask[k:k + len(x)], acc), acc)

    return lane_sum(term, acc_fn, adjacent=True)


def klt_sad(p1, p2):
    """klt.h:139-149"""
    p1 = np.asarray(p1, np.float32).ravel()
    p2 = np.asarray(p2, np.float32).ravel()
    d = p1 - p2
    term = d * d
    keep = (p1 != 0) & (p2 != 0)
    mask = _mask()

    def acc_fn(acc, x, k):
        kk = keep[k:k + len(x)]
        return np.where(kk, fma(x, mask[k:k + len(x)], acc), acc)

    return lane_sum(term, acc_fn)


def brute_sad(p1, m1, q1, p2, m2, q2):
    """brute.h:82-94"""
    p1 = np.asarray(p1, np.float32).ravel()
    p2 = np.asarray(p2, np.float32).ravel()
    with np.errstate(all="ignore"):
        alpha = f32(np.sqrt(f32(q1) / f32(q2)))
        beta = f32(f32(m1) - alpha * f32(m2))
        diff = fma(-p2, alpha, p1) - beta
    keep = (p1 != 0) & (p2 != 0)

    def acc_fn(acc, x, k):
        kk = keep[k:k + len(x)]
        return np.where(kk, fma(x, x, acc), acc)

    return lane_sum(diff, acc_fn)


# ------------------------------------------------------------------ Newton machinery

def _fd(s, h, central):
    s = [float(v) for v in s]
    if central:  # hessian.h:163-169
        sad0, sadn1x, sadn1y, sadp1x, sadp1y, sadxy = s
        dx = 0.5 * (sadp1x - sadn1x) / h
        dy = 0.5 * (sadp1y - sadn1y) / h
        dxx = ((sadp1x - sad0) / h - (sad0 - sadn1x) / h) / h
        dyy = ((sadp1y - sad0) / h - (sad0 - sadn1y) / h) / h
        dxy = ((sadxy - sadp1y) / h - (sadp1x - sad0) / h) / h
        dyx = ((sadxy - sadp1x) / h - (sadp1y - sad0) / h) / h
    else:  # klt.h:197-203
        sad0, sadx, sady, sadxx, sadyy, sadxy = s
        dx = (sadx - sad0) / h
        dy = (sady - sad0) / h
        dxx = ((sadxx - sadx) / h - (sadx - sad0) / h) / h
        dyy = ((sadyy - sady) / h - (sady - sad0) / h) / h
        dxy = ((sadxy - sady) / h - (sadx - sad0) / h) / h
        dyx = ((sadxy - sadx) / h - (sady - sad0) / h) / h
    return [f32(v) for v in (dx, dy, dxx, dxy, dyx, dyy)]


def hes_brute_hessian(img, patch, pm, pq, x, y):
    """hessian.h:147-172"""
    h = 0.02
    x, y = f32(x), f32(y)
    xm, xp = f32(float(x) - h), f32(float(x) + h)
    ym, yp = f32(float(y) - h), f32(float(y) + h)
    pts = [(x, y), (xm, y), (x, ym), (xp, y), (x, yp), (xp, yp)]
    s = []
    for (qx, qy) in pts:
        d, m, q = hes_get_patch(img, qx, qy)
        s.append(hes_score(patch, pm, pq, d, m, q))
    return s[0], _fd(s, h, True)


def newton_step(d6):
    """hessian.h:209-227 (Eigen 2x2 inverse closed form, then the normalisation quirk)"""
    with np.errstate(all="ignore"):
        dx_, dy_, dxx, dxy, dyx, dyy = [np.float64(v) for v in d6]
        det = dxx * dyy - dyx * dxy
        invdet = np.float64(1.0) / det
        i00, i10, i01, i11 = dyy * invdet, -dyx * invdet, -dxy * invdet, dxx * invdet
        j0 = i00 * dx_ + i01 * dy_
        j1 = i10 * dx_ + i11 * dy_
        dx, dy = f32(-j0), f32(-j1)
        if f32(f32(dx * dx) + f32(dy * dy)) > 1:
            dx = f32(dx / f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))))
            dy = f32(dy / f32(np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))))
    return dx, dy


def _clamp1(v):
    # max(-1.f, min(1.f, v)) with std::min/std::max NaN behaviour
    m = v if v < f32(1) else f32(1)
    return m if f32(-1) < m else f32(-1)


def hes_track(img, patch, pm, pq, thr, maxit, x, y, cnt):
    """hessian.h:185-241"""
    x, y = f32(x), f32(y)
    margin = f32(0.01)
    h_, w_ = img.shape
    thr = f32(thr)
    for _ in range(maxit):
        if x < margin or y < margin or f32(x + margin) > f32(w_) or f32(y + margin) > f32(h_):
            return OUT_OF_BOUNDS, x, y
        _, d6 = hes_brute_hessian(img, patch, pm, pq, x, y)
        cnt[0] += 1
        cnt[1] += 6
        dx, dy = newton_step(d6)
        x = f32(x + _clamp1(dx))
        y = f32(y + _clamp1(dy))
        if abs(dx) < thr and abs(dy) < thr:
            break
    return OK, x, y


def hes_track_feature(tmpl_pyr, tx, ty, search_pyr, levels, thr, maxit, x, y, cnt):
    """GetPatches (hessian.h:175-183) + TrackFeature (hessian.h:243-264)"""
    lv = min(len(tmpl_pyr), levels)
    patches = []
    qx, qy = f32(tx), f32(ty)
    for i in range(lv):
        patches.append(hes_get_patch(tmpl_pyr[i], qx, qy))
        qx, qy = f32(qx * f32(0.5)), f32(qy * f32(0.5))
    cnt[1] += lv
    lvls = min(len(search_pyr), lv)
    scale = f32(1. / (1 << (lvls - 1)))
    px, py = f32(f32(x) * scale), f32(f32(y) * scale)
    for i in range(lvls - 1, 0, -1):
        st, px, py = hes_track(search_pyr[i], *patches[i], thr, maxit, px, py, cnt)
        if st != OK:
            return st, f32(x), f32(y)
        px, py = f32(px * f32(2)), f32(py * f32(2))
    st, px, py = hes_track(search_pyr[0], *patches[0], thr, maxit, px, py, cnt)
    if st != OK:
        return st, f32(x), f32(y)
    return OK, px, py


def hes_track_fb(pfrom, pto, from_xy, seed_xy, levels, thr=0.001, maxit=10, fb_max=0.3):
    """matcher.cpp:173-206"""
    from_xy = np.asarray(from_xy, np.float32).reshape(-1, 2)
    seed_xy = np.asarray(seed_xy, np.float32).reshape(-1, 2)
    n = len(from_xy)
    levels = np.broadcast_to(np.asarray(levels, np.int32), (n,))
    out = dict(to_xy=np.empty((n, 2), np.float32), back_xy=np.empty((n, 2), np.float32),
               status_fwd=np.empty(n, np.int32), status_bwd=np.empty(n, np.int32),
               accepted=np.empty(n, np.uint8))
    cnt = [0, 0]
    for i in range(n):
        fx, fy = from_xy[i]
        s1, tx, ty = hes_track_feature(pfrom, fx, fy, pto, int(levels[i]), thr, maxit, seed_xy[i, 0], seed_xy[i, 1], cnt)
        s2, bx, by = hes_track_feature(pto, tx, ty, pfrom, int(levels[i]), thr, maxit, fx, fy, cnt)
        ok = not (s1 or s2)
        if ok:
            ddx, ddy = f32(fx - bx), f32(fy - by)
            if math.sqrt(float(ddx) * float(ddx) + float(ddy) * float(ddy)) > float(fb_max):
                ok = False
        out["to_xy"][i] = (tx, ty)
        out["back_xy"][i] = (bx, by)
        out["status_fwd"][i] = s1
        out["status_bwd"][i] = s2
        out["accepted"][i] = ok
    out["newton_steps"], out["patches"] = cnt
    return out


# ------------------------------------------------------------------ P4 external oracle

def hamming_knn2_cv2(q_bytes, t_bytes):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) -> idx[nq,2], dist[nq,2]"""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    res = bf.knnMatch(np.ascontiguousarray(q_bytes), np.ascontiguousarray(t_bytes), k=2)
    nq = len(res)
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), 257, np.int32)
    for i, ms in enumerate(res):
        for k, m in enumerate(ms[:2]):
            idx[i, k] = m.trainIdx
            dist[i, k] = int(m.distance)
    return idx, dist


# ------------------------------------------------------------------ P2 tier-0: KLTTracker (klt.h) on cv2 planes

def klt_get_patch(level_planes, x, y):
    """klt.h:59-96: three full 13x13 cv2.getRectSubPix calls (image, gradx, grady), no clipping."""
    img, gx, gy = level_planes
    c = (float(f32(x)), float(f32(y)))
    data = cv2.getRectSubPix(img, (N, N), c)
    m, q = patch_stats(data)
    return data.ravel(), cv2.getRectSubPix(gx, (N, N), c).ravel(), cv2.getRectSubPix(gy, (N, N), c).ravel(), m, q


def _lane_sums(terms, keep):
    """Declared reduction of several masked fused accumulations at once: terms[j] is a list of (a, b) factor
    arrays, accumulated per pixel as acc = fma(a, b, acc) where keep is set."""
    out = []
    for (a, b) in terms:
        a = np.asarray(a, np.float32)
        b = np.asarray(b, np.float32)

        def acc_fn(acc, x, k, a=a, b=b):
            kk = keep[k:k + len(x)]
            return np.where(kk, fma(a[k:k + len(x)], b[k:k + len(x)], acc), acc)

        out.append(lane_sum(a, acc_fn))
    return out


def klt_system(tmpl, cand):
    """klt.h:286-343 for one (template patch, candidate patch) pair: the 24 numbers
    A(4) B(4) C(4) RS(2) VW(2) U(4) e(2) d(2); tmpl / cand = (data, gradx, grady, mean, sumsq)."""
    I, Igx, Igy, Im, Iq = tmpl
    J0, Jgx, Jgy, Jm, Jq = cand
    mask = _mask()
    with np.errstate(all="ignore"):
        alpha = f32(np.sqrt(f32(Iq) / f32(Jq)))             # :289
        beta = f32(f32(Im) - alpha * f32(Jm))               # :290
        keep = (J0 != 0) & (I != 0)                         # :303
        Jv = fma(J0, alpha, beta)                           # :308
        gJ0, gJ1 = (Jgx * alpha).astype(np.float32), (Jgy * alpha).astype(np.float32)  # :313
        diff = ((I - Jv) * mask).astype(np.float32)         # :319
        pairs = [(Igx * Igx, mask), (Igx * Igy, mask), (Igy * Igx, mask), (Igy * Igy, mask),          # A :316
                 (Igx * gJ0, mask), (Igx * gJ1, mask), (Igy * gJ0, mask), (Igy * gJ1, mask),          # B :317
                 (gJ0 * gJ0, mask), (gJ0 * gJ1, mask), (gJ1 * gJ0, mask), (gJ1 * gJ1, mask),          # C :318
                 (diff, Igx), (diff, Igy), (diff, gJ0), (diff, gJ1)]                                  # RS, VW :320-321
        s = _lane_sums(pairs, keep)
        A, B, C, RS, VW = s[0:4], s[4:8], s[8:12], s[12:14], s[14:16]
        lam = f32(.0001)                                    # :274
        t00, t01, t10, t11 = B[0], B[2], B[1], B[3]         # B transposed (:326), Eigen's 2x2 closed-form inverse
        det = f32(f32(t00 * t11) - f32(t10 * t01))
        inv = f32(f32(1) / det)
        D = [f32(t11 * inv), f32(-t01 * inv), f32(-t10 * inv), f32(t00 * inv)]
        Al = [f32(A[0] + lam), A[1], A[2], f32(A[3] + lam)]
        M = [f32(f32(Al[0] * D[0]) + f32(Al[1] * D[2])), f32(f32(Al[0] * D[1]) + f32(Al[1] * D[3])),
             f32(f32(Al[2] * D[0]) + f32(Al[3] * D[2])), f32(f32(Al[2] * D[1]) + f32(Al[3] * D[3]))]
        h = f32(0.5)
        U = [f32(f32(f32(M[0] * C[0]) + f32(M[1] * C[2])) - f32(h * B[0])), f32(f32(f32(M[0] * C[1]) + f32(M[1] * C[3])) - f32(h * B[1])),
             f32(f32(f32(M[2] * C[0]) + f32(M[3] * C[2])) - f32(h * B[2])), f32(f32(f32(M[2] * C[1]) + f32(M[3] * C[3])) - f32(h * B[3]))]  # :330
        e = [f32(f32(f32(M[0] * VW[0]) + f32(M[1] * VW[1])) - f32(h * RS[0])),
             f32(f32(f32(M[2] * VW[0]) + f32(M[3] * VW[1])) - f32(h * RS[1]))]                                                               # :331
        u00, u01, u10, u11, e0, e1 = U[0], U[1], U[2], U[3], e[0], e[1]
        if abs(u10) > abs(u00):                             # U.lu().solve(e) (:343): 2x2 partial-pivot LU
            u00, u10, u01, u11, e0, e1 = u10, u00, u11, u01, e1, e0
        l = f32(u10 / u00)
        w11 = f32(u11 - f32(l * u01))
        y1 = f32(e1 - f32(l * e0))
        d1 = f32(y1 / w11)
        d0 = f32(f32(e0 - f32(u01 * d1)) / u00)
    return np.float32(list(A) + list(B) + list(C) + list(RS) + list(VW) + U + e + [d0, d1])


def klt_brute_hessian(level_planes, patch, x, y):
    """klt.h:181-204: forward differences, h = 0.01 (double)."""
    h = 0.01
    x, y = f32(x), f32(y)
    x1, x2 = f32(float(x) + h), f32(float(x) + 2 * h)
    y1, y2 = f32(float(y) + h), f32(float(y) + 2 * h)
    pts = [(x, y), (x1, y), (x, y1), (x2, y), (x, y2), (x1, y1)]
    s = [klt_sad(patch, cv2.getRectSubPix(level_planes[0], (N, N), (float(qx), float(qy)))) for (qx, qy) in pts]
    return _fd(s, h, False)


def klt_track(level_planes, tmpl, thr, maxit, x, y, trace):
    """klt.h:258-401 as written: the symmetric-KLT step is computed (and recorded in `trace`), then replaced by
    the finite-difference Newton step (klt.h:355-380)."""
    x, y = f32(x), f32(y)
    margin = f32(0.1)
    h_, w_ = level_planes[0].shape
    for _ in range(maxit):
        if x < margin or y < margin or f32(x + margin) > f32(w_) or f32(y + margin) > f32(h_):
            return OUT_OF_BOUNDS, x, y
        sys24 = klt_system(tmpl, klt_get_patch(level_planes, x, y))
        d6 = klt_brute_hessian(level_planes, tmpl[0], x, y)
        if trace is not None:
            trace.append((float(x), float(y), sys24, np.float32(d6)))
        dx, dy = newton_step(d6)
        x = f32(x + _clamp1(dx))
        y = f32(y + _clamp1(dy))
        if float(abs(dx)) < float(f32(thr)) / 10. and float(abs(dy)) < float(f32(thr)) / 10.:   # :392 (double compare)
            break
    return OK, x, y


def klt_track_feature(tmpl_pyr, tx, ty, search_pyr, thr, maxit, x, y, traces=None):
    """klt.h:249-256 + :403-424: all levels of the stack, coarse threshold x50."""
    lvls = min(len(tmpl_pyr), len(search_pyr))
    patches = []
    qx, qy = f32(tx), f32(ty)
    for i in range(lvls):
        patches.append(klt_get_patch(tmpl_pyr[i], qx, qy))
        qx, qy = f32(qx * f32(0.5)), f32(qy * f32(0.5))
    scale = f32(1. / (1 << (lvls - 1)))
    px, py = f32(f32(x) * scale), f32(f32(y) * scale)
    for i in range(lvls - 1, -1, -1):
        tr = None
        if traces is not None:
            tr = []
            traces.append((i, tr))
        st, px, py = klt_track(search_pyr[i], patches[i], f32(f32(thr) * f32(50)) if i > 0 else f32(thr), maxit, px, py, tr)
        if st != OK:
            return st, f32(x), f32(y)
        if i > 0:
            px, py = f32(px * f32(2)), f32(py * f32(2))
    return OK, px, py


def klt_track_fb(pfrom, pto, from_xy, thr=0.001, maxit=10, fb_max=0.3, traces=None):
    """matcher.cpp:173-206 driven with the KLT tracker (seed = from_pt)."""
    from_xy = np.asarray(from_xy, np.float32).reshape(-1, 2)
    n = len(from_xy)
    out = dict(to_xy=np.empty((n, 2), np.float32), back_xy=np.empty((n, 2), np.float32), status_fwd=np.empty(n, np.int32),
               status_bwd=np.empty(n, np.int32), accepted=np.empty(n, np.uint8))
    for i in range(n):
        fx, fy = from_xy[i]
        tr = [] if traces is not None else None
        s1, tx, ty = klt_track_feature(pfrom, fx, fy, pto, thr, maxit, fx, fy, tr)
        s2, bx, by = klt_track_feature(pto, tx, ty, pfrom, thr, maxit, fx, fy, None)
        if traces is not None:
            traces.append(tr)
        ok = not (s1 or s2)
        if ok:
            ddx, ddy = f32(fx - bx), f32(fy - by)
            if math.sqrt(float(ddx) * float(ddx) + float(ddy) * float(ddy)) > float(fb_max):
                ok = False
        out["to_xy"][i] = (tx, ty)
        out["back_xy"][i] = (bx, by)
        out["status_fwd"][i], out["status_bwd"][i], out["accepted"][i] = s1, s2, ok
    return out


# ------------------------------------------------------------------ P3 tier-0: BruteTracker (brute.h) on cv2 planes

def _float_offsets(window, res):
    """The offsets visited by `for (float x = -window; x <= window; x += res)` (brute.h:105-106), float32 counters."""
    out = []
    x, w, r = f32(-f32(window)), f32(window), f32(res)
    while x <= w:
        out.append(x)
        x = f32(x + r)
    return np.float32(out)


def _batched_stats_and_sad(tmpl, tm, tq, cand):
    """brute.h:34-57 statistics and brute.h:82-94 SAD for a batch of candidate patches [P,169], declared order,
    vectorised over the batch."""
    P = cand.shape[0]
    pad = np.zeros((P, 192), np.float32)
    pad[:, :LEN] = cand
    lanes = pad.reshape(P, 6, 32)
    s = np.zeros((P, 32), np.float32)
    q = np.zeros((P, 32), np.float32)
    for k in range(6):
        n = 32 if k < 5 else LEN - 160
        s[:, :n] = s[:, :n] + lanes[:, k, :n]
        q[:, :n] = fma(lanes[:, k, :n], lanes[:, k, :n], q[:, :n])

    def tree(v):
        v = v.copy()
        off = 16
        while off:
            v[:, :off] = v[:, :off] + v[:, off:2 * off]
            off >>= 1
        return v[:, 0]

    mean = (tree(s) / f32(LEN)).astype(np.float32)
    sumsq = (tree(q) / f32(LEN)).astype(np.float32)
    with np.errstate(all="ignore"):
        alpha = np.sqrt(f32(tq) / sumsq).astype(np.float32)
        beta = (f32(tm) - alpha * mean).astype(np.float32)
        tp = np.zeros(192, np.float32)
        tp[:LEN] = tmpl
        tl = tp.reshape(6, 32)
        acc = np.zeros((P, 32), np.float32)
        for k in range(6):
            n = 32 if k < 5 else LEN - 160
            diff = (fma(-lanes[:, k, :n], alpha[:, None], tl[k, :n][None, :]) - beta[:, None]).astype(np.float32)
            keep = (lanes[:, k, :n] != 0) & (tl[k, :n][None, :] != 0)
            acc[:, :n] = np.where(keep, fma(diff, diff, acc[:, :n]), acc[:, :n])
    return tree(acc)


def brute_search_best(img, tmpl, tm, tq, window, res, px, py):
    """brute.h:96-117: x outer, y inner, `if (sad > best) continue` -> the LAST minimum wins."""
    offs = _float_offsets(window, res)
    px, py = f32(px), f32(py)
    best, bx, by = f32(1e6), px, py
    for ox in offs:
        cx = f32(px + ox)
        cys = (py + offs).astype(np.float32)
        cand = np.stack([cv2.getRectSubPix(img, (N, N), (float(cx), float(cy))).ravel() for cy in cys])
        sads = _batched_stats_and_sad(tmpl, tm, tq, cand)
        for j in range(len(cys)):   # sequential rule; NaN compares false, i.e. a NaN score is accepted like the reference
            if sads[j] > best:
                continue
            bx, by, best = cx, cys[j], sads[j]
    return best, bx, by, len(offs) * len(offs)


BRUTE_COARSE = [(3, 1), (1, 0.33333)]                                         # brute.h:147-148
BRUTE_FINE = [(3, 1), (1, 0.3333), (0.4, 0.1), (0.2, 0.025), (8, 0.01)]       # brute.h:154-158, as written


def brute_track_feature(tmpl_pyr, tx, ty, search_pyr, x, y, coarse=BRUTE_COARSE, fine=BRUTE_FINE, trace=None):
    """brute.h:129-164 with the schedule as written (incl. the (8, 0.01) pass)."""
    lvls = min(len(tmpl_pyr), len(search_pyr))
    margin = f32(13)
    h0, w0 = search_pyr[0].shape
    x, y = f32(x), f32(y)
    if x < margin or y < margin or f32(x + margin) > f32(w0) or f32(y + margin) > f32(h0):
        return OUT_OF_BOUNDS, x, y, f32(0)
    patches = []
    qx, qy = f32(tx), f32(ty)
    for i in range(lvls):
        patches.append(full_patch(tmpl_pyr[i], qx, qy))
        qx, qy = f32(qx * f32(0.5)), f32(qy * f32(0.5))
    scale = f32(1. / (1 << (lvls - 1)))
    px, py = f32(x * scale), f32(y * scale)
    sad = f32(0)
    for i in range(lvls - 1, -1, -1):
        d, m, q = patches[i]
        for (win, res) in (coarse if i > 0 else fine):
            sad, px, py, npos = brute_search_best(search_pyr[i], d.ravel(), m, q, win, res, px, py)
            if trace is not None:
                trace.append((i, float(win), float(res), float(px), float(py), float(sad), npos))
        if sad > 100:
            return OUT_OF_BOUNDS, x, y, sad
        if i > 0:
            px, py = f32(px * f32(2)), f32(py * f32(2))
    return OK, px, py, sad
