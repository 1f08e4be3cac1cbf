"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
The product package (slam-robot_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OK, SMALL_DET, OUT_OF_BOUNDS = 0, 1, 2
FLAVOR_HESSIAN, FLAVOR_KLT, FLAVOR_BRUTE = 0, 1, 2

# brute.h:147-158 search schedules ({window, res} pairs)
BRUTE_COARSE = np.array([3, 1, 1, 0.33333], dtype=np.float32)
# BRUTE_FINE is the level-0 schedule as written (brute.h:154-158, incl. the (8, 0.01) pass); _FAST drops that pass
BRUTE_FINE = np.array([3, 1, 1, 0.3333, 0.4, 0.1, 0.2, 0.025, 8, 0.01], dtype=np.float32)
BRUTE_FINE_FAST = np.array([3, 1, 1, 0.3333, 0.4, 0.1, 0.2, 0.025], dtype=np.float32)


class Counters(C.Structure):
    _fields_ = [("newton_steps", C.c_int64), ("patches", C.c_int64)]


def build(fast=False, native=False):
    """Compile the oracle with the system gcc. Returns the path of the .so."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    out = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "oracle.c")
    if (not native and os.path.exists(out)
            and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle.h")))):
        return out
    subprocess.check_call(["make", "-C", _HERE, "-B", name], stdout=subprocess.DEVNULL)
    return out


_libs = {}


def lib(fast=False, native=False):
    key = bool(fast)
    if key in _libs:
        return _libs[key]
    path = build(fast=fast, native=native)
    L = C.CDLL(path)
    f32p, i32p, u8p, u32p, i64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_int64))
    vp = C.c_void_p
    L.orc_mask13.argtypes = [f32p]
    L.orc_gray_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p]
    L.orc_gauss5_sigma.argtypes = [f32p, C.c_int, C.c_int, C.c_double, f32p]
    L.orc_pyrdown.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.orc_scharr.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p]
    L.orc_rect_subpix.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p, C.c_int]
    L.orc_pyr_build.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int]
    L.orc_pyr_build.restype = vp
    L.orc_pyr_from_planes.argtypes = [C.c_int, C.c_int, i32p, i32p, C.POINTER(f32p)]
    L.orc_pyr_from_planes.restype = vp
    L.orc_pyr_free.argtypes = [vp]
    for fn in (L.orc_pyr_depth,):
        fn.argtypes = [vp]
        fn.restype = C.c_int
    for fn in (L.orc_pyr_w, L.orc_pyr_h):
        fn.argtypes = [vp, C.c_int]
        fn.restype = C.c_int
    L.orc_pyr_plane.argtypes = [vp, C.c_int, C.c_int]
    L.orc_pyr_plane.restype = f32p
    L.orc_hes_get_patch.argtypes = [vp, C.c_int, C.c_float, C.c_float, f32p, f32p, f32p]
    L.orc_hes_score.argtypes = [f32p, C.c_float, C.c_float, f32p, C.c_float, C.c_float]
    L.orc_hes_score.restype = C.c_float
    L.orc_hes_brute_hessian.argtypes = [vp, C.c_int, f32p, C.c_float, C.c_float, C.c_float, C.c_float, f32p]
    L.orc_hes_brute_hessian.restype = C.c_float
    L.orc_hes_track_feature.argtypes = [vp, C.c_float, C.c_float, vp, C.c_int, C.c_float, C.c_int, f32p, f32p,
                                        C.POINTER(Counters)]
    L.orc_hes_track_feature.restype = C.c_int
    L.orc_hes_track_fb.argtypes = [vp, vp, C.c_int, f32p, f32p, i32p, C.c_float, C.c_int, C.c_double, f32p, i32p,
                                   i32p, u8p, C.POINTER(Counters), C.c_int]
    L.orc_hes_track_fb.restype = C.c_int
    L.orc_klt_system.argtypes = [vp, C.c_float, C.c_float, vp, C.c_int, C.c_float, C.c_float, f32p]
    L.orc_klt_track_feature.argtypes = [vp, C.c_float, C.c_float, vp, C.c_float, C.c_int, f32p, f32p,
                                        C.POINTER(Counters)]
    L.orc_klt_track_feature.restype = C.c_int
    L.orc_klt_track_fb.argtypes = [vp, vp, C.c_int, f32p, f32p, C.c_float, C.c_int, C.c_double, f32p, i32p, i32p,
                                   u8p, C.POINTER(Counters), C.c_int]
    L.orc_klt_track_fb.restype = C.c_int
    L.orc_brute_search_best.argtypes = [vp, C.c_int, f32p, C.c_float, C.c_float, C.c_float, C.c_float, f32p,
                                        f32p, i64p]
    L.orc_brute_search_best.restype = C.c_float
    L.orc_brute_track.argtypes = [vp, vp, C.c_int, f32p, f32p, f32p, C.c_int, f32p, C.c_int, i32p, f32p, i64p,
                                  C.c_int]
    L.orc_brute_track.restype = C.c_int
    f64p = C.POINTER(C.c_double)
    L.orc_project.argtypes = [f64p, f64p, f64p, f64p, f64p]
    L.orc_seed_features.argtypes = [C.c_int, f64p, f64p, f64p, f64p, f64p, f32p, C.c_int, C.c_int, f32p, i32p, u8p]
    L.orc_yuyv_to_bgr.argtypes = [u8p, C.c_size_t, u8p]
    L.orc_min_eigen_val.argtypes = [u8p, C.c_int, C.c_int, f32p]
    L.orc_good_features.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double, f32p, f32p, f32p]
    L.orc_hamming256_top2.argtypes = [u32p, C.c_int, u32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, u8p,
                                      C.c_int]
    L.orc_num_threads.restype = C.c_int
    _libs[key] = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def mask13():
    m = np.empty(169, np.float32)
    lib().orc_mask13(_p(m, C.c_float))
    return m


def gray_u8(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    g = np.empty((h, w), np.uint8)
    lib().orc_gray_u8(_p(bgr, C.c_uint8), w, h, bgr.strides[0], _p(g, C.c_uint8))
    return g


def gauss5(img, sigma):
    img = _f32(img)
    out = np.empty_like(img)
    lib().orc_gauss5_sigma(_p(img, C.c_float), img.shape[1], img.shape[0], float(sigma), _p(out, C.c_float))
    return out


def pyrdown(img):
    img = _f32(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyrdown(_p(img, C.c_float), w, h, _p(out, C.c_float))
    return out


def scharr(img):
    img = _f32(img)
    gx = np.empty_like(img)
    gy = np.empty_like(img)
    lib().orc_scharr(_p(img, C.c_float), img.shape[1], img.shape[0], _p(gx, C.c_float), _p(gy, C.c_float))
    return gx, gy


def rect_subpix(img, n, m, cx, cy):
    img = _f32(img)
    out = np.empty((m, n), np.float32)
    lib().orc_rect_subpix(_p(img, C.c_float), img.shape[1], img.shape[0], n, m, C.c_float(cx), C.c_float(cy),
                          _p(out, C.c_float), n)
    return out


class Pyramid:
    """orc_pyr handle (MakePyramid of hessian.h / klt.h / brute.h)."""

    def __init__(self, bgr, depth, flavor=FLAVOR_HESSIAN, fast=False):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        assert bgr.ndim == 3 and bgr.shape[2] == 3
        self._L = lib(fast=fast)
        h, w, _ = bgr.shape
        self.h = self._L.orc_pyr_build(_p(bgr, C.c_uint8), w, h, bgr.strides[0], depth, flavor)
        if not self.h:
            raise ValueError("orc_pyr_build failed")
        self.depth = depth
        self.flavor = flavor

    @classmethod
    def from_planes(cls, planes, flavor=FLAVOR_HESSIAN, fast=False):
        """planes: list over levels of a 2-D array (image only) or a tuple (img, gx, gy)."""
        self = cls.__new__(cls)
        self._L = lib(fast=fast)
        depth = len(planes)
        ws = (C.c_int32 * depth)()
        hs = (C.c_int32 * depth)()
        ptrs = (C.POINTER(C.c_float) * (3 * depth))()
        keep = []
        for i, lvl in enumerate(planes):
            trio = lvl if isinstance(lvl, (tuple, list)) else (lvl, None, None)
            hs[i], ws[i] = trio[0].shape
            for k, a in enumerate(trio):
                if a is None:
                    continue
                a = _f32(a)
                keep.append(a)
                ptrs[3 * i + k] = _p(a, C.c_float)
        self.h = self._L.orc_pyr_from_planes(depth, flavor, ws, hs, ptrs)
        self.depth, self.flavor = depth, flavor
        return self

    def __del__(self):
        if getattr(self, "h", None):
            self._L.orc_pyr_free(self.h)
            self.h = None

    def size(self, level):
        return self._L.orc_pyr_w(self.h, level), self._L.orc_pyr_h(self.h, level)

    def plane(self, level, which=0):
        w, h = self.size(level)
        ptr = self._L.orc_pyr_plane(self.h, level, which)
        return np.ctypeslib.as_array(ptr, shape=(h, w)).copy()


def hes_get_patch(pyr, level, x, y):
    d = np.empty(169, np.float32)
    mean = C.c_float()
    sumsq = C.c_float()
    pyr._L.orc_hes_get_patch(pyr.h, level, C.c_float(x), C.c_float(y), _p(d, C.c_float), C.byref(mean),
                             C.byref(sumsq))
    return d.reshape(13, 13), np.float32(mean.value), np.float32(sumsq.value)


def hes_score(p1, m1, q1, p2, m2, q2):
    p1 = _f32(p1).ravel()
    p2 = _f32(p2).ravel()
    return np.float32(lib().orc_hes_score(_p(p1, C.c_float), C.c_float(m1), C.c_float(q1), _p(p2, C.c_float),
                                         C.c_float(m2), C.c_float(q2)))


def hes_brute_hessian(pyr, level, patch, mean, sumsq, x, y):
    patch = _f32(patch).ravel()
    out = np.empty(6, np.float32)
    s0 = pyr._L.orc_hes_brute_hessian(pyr.h, level, _p(patch, C.c_float), C.c_float(mean), C.c_float(sumsq),
                                      C.c_float(x), C.c_float(y), _p(out, C.c_float))
    return np.float32(s0), out


def hes_track_fb(pfrom, pto, from_xy, seed_xy, levels, thr=0.001, maxit=10, fb_max=0.3, nthreads=0):
    """matcher.cpp:173-206 for n features. Returns dict of arrays + counters."""
    from_xy = _f32(from_xy).reshape(-1, 2)
    n = from_xy.shape[0]
    to_xy = _f32(seed_xy).reshape(-1, 2).copy()
    levels = np.ascontiguousarray(np.broadcast_to(np.asarray(levels, np.int32), (n,)))
    back = np.empty((n, 2), np.float32)
    s1 = np.empty(n, np.int32)
    s2 = np.empty(n, np.int32)
    acc = np.empty(n, np.uint8)
    cnt = Counters(0, 0)
    pfrom._L.orc_hes_track_fb(pfrom.h, pto.h, n, _p(from_xy, C.c_float), _p(to_xy, C.c_float),
                              _p(levels, C.c_int32), C.c_float(thr), maxit, C.c_double(fb_max),
                              _p(back, C.c_float), _p(s1, C.c_int32), _p(s2, C.c_int32), _p(acc, C.c_uint8),
                              C.byref(cnt), nthreads)
    return dict(to_xy=to_xy, back_xy=back, status_fwd=s1, status_bwd=s2, accepted=acc,
                newton_steps=cnt.newton_steps, patches=cnt.patches)


def klt_system(pfrom, tx, ty, pto, level, x, y):
    out = np.empty(24, np.float32)
    pfrom._L.orc_klt_system(pfrom.h, C.c_float(tx), C.c_float(ty), pto.h, level, C.c_float(x), C.c_float(y),
                            _p(out, C.c_float))
    return out


def klt_track_fb(pfrom, pto, from_xy, seed_xy, thr=0.001, maxit=10, fb_max=0.3, nthreads=0):
    from_xy = _f32(from_xy).reshape(-1, 2)
    n = from_xy.shape[0]
    to_xy = _f32(seed_xy).reshape(-1, 2).copy()
    back = np.empty((n, 2), np.float32)
    s1 = np.empty(n, np.int32)
    s2 = np.empty(n, np.int32)
    acc = np.empty(n, np.uint8)
    cnt = Counters(0, 0)
    pfrom._L.orc_klt_track_fb(pfrom.h, pto.h, n, _p(from_xy, C.c_float), _p(to_xy, C.c_float), C.c_float(thr),
                              maxit, C.c_double(fb_max), _p(back, C.c_float), _p(s1, C.c_int32),
                              _p(s2, C.c_int32), _p(acc, C.c_uint8), C.byref(cnt), nthreads)
    return dict(to_xy=to_xy, back_xy=back, status_fwd=s1, status_bwd=s2, accepted=acc,
                newton_steps=cnt.newton_steps, patches=cnt.patches)


def brute_track(pfrom, pto, from_xy, seed_xy, coarse=BRUTE_COARSE, fine=BRUTE_FINE, nthreads=0):
    from_xy = _f32(from_xy).reshape(-1, 2)
    n = from_xy.shape[0]
    to_xy = _f32(seed_xy).reshape(-1, 2).copy()
    st = np.empty(n, np.int32)
    sad = np.zeros(n, np.float32)
    npos = C.c_int64(0)
    coarse = _f32(coarse)
    fine = _f32(fine)
    pfrom._L.orc_brute_track(pfrom.h, pto.h, n, _p(from_xy, C.c_float), _p(to_xy, C.c_float),
                             _p(coarse, C.c_float), len(coarse) // 2, _p(fine, C.c_float), len(fine) // 2,
                             _p(st, C.c_int32), _p(sad, C.c_float), C.byref(npos), nthreads)
    return dict(to_xy=to_xy, status=st, best_sad=sad, positions=npos.value)


def hamming256_top2(q, t, ratio_num=4, ratio_den=5, max_dist=256, nthreads=0, fast=False):
    q = np.ascontiguousarray(q).view(np.uint32).reshape(-1, 8)
    t = np.ascontiguousarray(t).view(np.uint32).reshape(-1, 8)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.int32)
    ok = np.empty(nq, np.uint8)
    lib(fast=fast).orc_hamming256_top2(_p(q, C.c_uint32), nq, _p(t, C.c_uint32), nt, ratio_num, ratio_den,
                                       max_dist, _p(idx, C.c_int32), _p(dist, C.c_int32), _p(ok, C.c_uint8),
                                       nthreads)
    return idx, dist, ok


def num_threads():
    return lib().orc_num_threads()


def min_eigen_val(gray):
    """cv::cornerMinEigenVal(gray8, blockSize 3, ksize 3)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.empty((h, w), np.float32)
    lib().orc_min_eigen_val(_p(gray, C.c_uint8), w, h, _p(out, C.c_float))
    return out


def good_features(bgr, max_corners=120, quality=0.01, min_distance=20.0, want_eig=False):
    """cv::goodFeaturesToTrack on the RGB2GRAY image of a BGR frame (matcher.cpp:313 + :123-130)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    xy = np.zeros((max_corners, 2), np.float32)
    eig = np.empty((h, w), np.float32)
    mx = np.zeros(1, np.float32)
    n = lib().orc_good_features(_p(bgr, C.c_uint8), w, h, bgr.strides[0], max_corners, float(quality), float(min_distance),
                                _p(xy, C.c_float), _p(eig, C.c_float), _p(mx, C.c_float))
    return (xy[:n].copy(), eig, float(mx[0])) if want_eig else xy[:n].copy()


def seed_features(points4, uncertainty, rot, trans, k, from_xy, cols, rows):
    """matcher.cpp:224-245 for n features: (seed_xy, levels, go)."""
    pts = np.ascontiguousarray(points4, np.float64).reshape(-1, 4)
    n = len(pts)
    unc = np.ascontiguousarray(uncertainty, np.float64)
    rot, trans, k = (np.ascontiguousarray(a, np.float64) for a in (rot, trans, k))
    fxy = _f32(from_xy).reshape(-1, 2)
    seed = np.empty((n, 2), np.float32)
    lv = np.empty(n, np.int32)
    go = np.empty(n, np.uint8)
    lib().orc_seed_features(n, _p(pts, C.c_double), _p(unc, C.c_double), _p(rot, C.c_double), _p(trans, C.c_double),
                            _p(k, C.c_double), _p(fxy, C.c_float), int(cols), int(rows), _p(seed, C.c_float),
                            _p(lv, C.c_int32), _p(go, C.c_uint8))
    return seed, lv, go


def yuyv_to_bgr(yuyv):
    """video.cpp:187-223 on a flat uint8 YUYV buffer (2 bytes per pixel) -> flat BGR (3 bytes per pixel)."""
    yuyv = np.ascontiguousarray(yuyv, np.uint8).ravel()
    out = np.empty(yuyv.size // 4 * 6, np.uint8)
    lib().orc_yuyv_to_bgr(_p(yuyv, C.c_uint8), yuyv.size, _p(out, C.c_uint8))
    return out
