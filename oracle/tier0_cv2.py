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


def lane_sum(vals, acc_fn):
    """Declared reduction: pixel i -> lane i%32, sequential within a lane, then tree32."""
    vals = np.asarray(vals, np.float32)
    lanes = np.zeros(32, np.float32)
    for k in range(0, LEN, 32):
        chunk = vals[k:k + 32]
        lanes[:len(chunk)] = acc_fn(lanes[:len(chunk)], chunk, k)
    return tree32(lanes)


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

def patch_stats(data):
    d = np.asarray(data, np.float32).ravel()
    s = lane_sum(d, lambda acc, x, k: acc + x)
    q = lane_sum(d, lambda acc, x, k: fma(x, x, acc))
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
    m, q = patch_stats(data)
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

    return lane_sum(term, acc_fn)


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
            if math.sqrt(float(ddx) * float(ddx) + float(ddy) * float(ddy)) > float(f32(fb_max)):
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
