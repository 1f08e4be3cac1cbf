"""slam-robot_b200 -- B200-native visual front-end of ywrt/slam-robot.

This package is a thin ctypes binding of the C ABI in include/slamfe.h (libslamfe.so, built
from csrc/*.cu for sm_100a) plus the host-side batching helpers used by tests and bench.py.
PyTorch appears only as plumbing (device buffers, streams, torch.distributed).  There is no
CPU fallback: if the CUDA library is missing or no GPU is present, every compute call raises.

The directory name is not a Python identifier; import it with
    importlib.import_module("slam-robot_b200")      or      import slam_robot_b200  (alias module)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_CSRC, "libslamfe.so")

OK, SMALL_DET, OUT_OF_BOUNDS = 0, 1, 2          # hessian.h:48-52
HESSIAN, KLT, BRUTE = 0, 1, 2                   # pyramid flavours
MAX_LEVELS = 12

# brute.h:147-148 / :154-158 search schedules ({window,res} pairs), AS WRITTEN: the last level-0 pass (8, 0.01)
# -- 1600 x 1600 positions per feature under the reference's float loop counters -- is live code: it sets the final
# position and the `sad` of the `sad > 100` gate (brute.h:158-159).  BRUTE_FINE_FAST drops it (an explicitly named
# variant, 2.5 M fewer positions per feature; NOT what BruteTracker::TrackFeature computes).
BRUTE_COARSE = np.array([3, 1, 1, 0.33333], dtype=np.float32)
BRUTE_FINE = np.array([3, 1, 1, 0.3333, 0.4, 0.1, 0.2, 0.025, 8, 0.01], dtype=np.float32)
BRUTE_FINE_FAST = np.array([3, 1, 1, 0.3333, 0.4, 0.1, 0.2, 0.025], dtype=np.float32)

EXPORTS = [
    "sfe_create", "sfe_destroy", "sfe_last_error", "sfe_set_stream", "sfe_sync", "sfe_get_mask", "sfe_host_alloc",
    "sfe_host_free", "sfe_launch_count", "sfe_pyr_create", "sfe_pyr_destroy", "sfe_pyr_level_size",
    "sfe_pyr_bytes_per_frame", "sfe_pyr_build", "sfe_pyr_build_dev", "sfe_pyr_download", "sfe_track_fb",
    "sfe_track_fb_dev", "sfe_track", "sfe_track_dev", "sfe_get_patches", "sfe_brute_hessian", "sfe_klt_track_fb", "sfe_klt_track_fb_dev",
    "sfe_klt_system", "sfe_brute_track", "sfe_brute_track_dev", "sfe_match_hamming256", "sfe_match_hamming256_dev",
    "sfe_match_hamming256_async", "sfe_hamming_impl", "sfe_replay_pairs", "sfe_replay_sequence", "sfe_replay_sequence_yuyv", "sfe_good_features",
    "sfe_good_features_dev",
    "sfe_seed_features", "sfe_seed_features_dev", "sfe_yuyv_to_bgr", "sfe_yuyv_to_bgr_dev",
    "sfe_shard_range", "sfe_bind_host_to_device", "sfe_dist_unique_id", "sfe_dist_init", "sfe_dist_attach", "sfe_dist_shutdown", "sfe_allgather_rows_dev",
    "sfe_match_hamming256_sharded_dev", "sfe_match_hamming256_sharded",
]


class SlamFEError(RuntimeError):
    pass


def sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(_HERE, "..", "include", "slamfe.h")]


def build(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libslamfe.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    out = subprocess.run([os.path.join(_CSRC, "build.sh")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise SlamFEError("nvcc build of libslamfe.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library. Fails loudly when it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    vpp = C.POINTER(C.c_void_p)
    L.sfe_create.argtypes = [i32, vpp]
    L.sfe_destroy.argtypes = [vp]
    L.sfe_destroy.restype = None
    L.sfe_last_error.argtypes = [vp]
    L.sfe_last_error.restype = C.c_char_p
    L.sfe_set_stream.argtypes = [vp, vp]
    L.sfe_sync.argtypes = [vp]
    L.sfe_get_mask.argtypes = [vp, vp]
    L.sfe_host_alloc.argtypes = [vp, sz, vpp]
    L.sfe_host_free.argtypes = [vp, vp]
    L.sfe_launch_count.argtypes = [vp]
    L.sfe_hamming_impl.argtypes = [i32]
    L.sfe_bind_host_to_device.argtypes = [i32]
    L.sfe_launch_count.restype = C.c_int64
    L.sfe_pyr_create.argtypes = [vp, i32, i32, i32, i32, i32, vpp]
    L.sfe_pyr_destroy.argtypes = [vp]
    L.sfe_pyr_destroy.restype = None
    L.sfe_pyr_level_size.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.sfe_pyr_bytes_per_frame.argtypes = [vp]
    L.sfe_pyr_bytes_per_frame.restype = C.c_int64
    L.sfe_pyr_build.argtypes = [vp, vp, vp, sz, sz, i32, i32]
    L.sfe_pyr_build_dev.argtypes = [vp, vp, vp, sz, sz, i32, i32]
    L.sfe_pyr_download.argtypes = [vp, vp, i32, i32, i32, vp]
    f64 = C.c_double   # fb_max: matcher.cpp:201 compares with the double literal 0.3
    trk = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, f32, i32, f64, vp, vp, vp, vp, vp]
    L.sfe_track_fb.argtypes = trk
    L.sfe_track_fb_dev.argtypes = trk
    one = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, f32, i32, vp, vp]
    L.sfe_track.argtypes = one
    L.sfe_track_dev.argtypes = one
    L.sfe_get_patches.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.sfe_brute_hessian.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp]
    klt = [vp, vp, i32, vp, i32, i32, i32, vp, vp, f32, i32, f64, vp, vp, vp, vp, vp]
    L.sfe_klt_track_fb.argtypes = klt
    L.sfe_klt_track_fb_dev.argtypes = klt
    L.sfe_klt_system.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp]
    L.sfe_brute_track.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, vp, i32, vp, vp, vp]
    L.sfe_brute_track_dev.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, vp, i32, vp, vp]
    ham = [vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.sfe_match_hamming256.argtypes = ham
    L.sfe_match_hamming256_dev.argtypes = ham
    L.sfe_match_hamming256_async.argtypes = ham
    L.sfe_good_features.argtypes = [vp, vp, i32, i32, sz, sz, i32, i32, C.c_double, C.c_double, vp, vp, vp]
    L.sfe_good_features_dev.argtypes = [vp, vp, i32, i32, sz, sz, i32, i32, C.c_double, C.c_double, vp, vp]
    seed = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.sfe_seed_features.argtypes = seed
    L.sfe_seed_features_dev.argtypes = seed
    L.sfe_yuyv_to_bgr.argtypes = [vp, vp, sz, vp]
    L.sfe_yuyv_to_bgr_dev.argtypes = [vp, vp, sz, vp]
    L.sfe_replay_pairs.argtypes = [vp, i32, i32, i32, i32, vp, vp, sz, sz, i32, vp, vp, vp, i32, f32, i32, f64, vp, vp, vp,
                                   vp, vp, i32]
    L.sfe_replay_sequence.argtypes = [vp, i32, i32, i32, i32, i32, vp, sz, sz, i32, vp, vp, vp, i32, f32, i32, f64, vp, vp, vp,
                                      vp, vp, i32]
    L.sfe_replay_sequence_yuyv.argtypes = L.sfe_replay_sequence.argtypes
    i64 = C.c_int64
    L.sfe_shard_range.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    L.sfe_dist_unique_id.argtypes = [vp]
    L.sfe_dist_init.argtypes = [vp, vp, i32, i32]
    L.sfe_dist_attach.argtypes = [vp, vp, i32, i32]
    L.sfe_dist_shutdown.argtypes = [vp]
    L.sfe_allgather_rows_dev.argtypes = [vp, vp, sz, i64, vp]
    shd = [vp, vp, i64, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.sfe_match_hamming256_sharded_dev.argtypes = shd
    L.sfe_match_hamming256_sharded.argtypes = shd
    _lib = L
    return L


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    """Raw address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr()
    return x.ctypes.data


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def bind_host_to_device(device):
    """sfe_bind_host_to_device: bind this process to the CPUs next to `device` (before allocating pinned buffers).
    Returns the number of CPUs bound to, 0 when nothing changed."""
    return int(lib().sfe_bind_host_to_device(int(device)))


class FrontEnd:
    """One sfe_ctx: one GPU, one stream.  Host-side mirror of the reference's FeatureTracker
    object (matcher.cpp:304) with batched entry points."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sfe_create(int(device), C.byref(h))
        if rc:
            raise SlamFEError("sfe_create(device=%d) failed with code %d: no usable CUDA device" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.L.sfe_host_free(self.h, p)
            self._pinned = []
            self.L.sfe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise SlamFEError("libslamfe error %d: %s" % (rc, self.L.sfe_last_error(self.h).decode()))

    def set_stream(self, cuda_stream_ptr):
        """None -> the context's own (non-blocking) stream; a cudaStream_t value -> that stream.  0 is CUDA's legacy
        default stream (what torch.cuda.current_stream().cuda_stream returns unless a stream context is active): it is
        passed as the explicit handle cudaStreamLegacy (0x1), because a NULL pointer means "own stream" in the C ABI."""
        if cuda_stream_ptr is None:
            ptr = 0
        else:
            ptr = int(cuda_stream_ptr) or 0x1
        self._chk(self.L.sfe_set_stream(self.h, C.c_void_p(ptr)))

    def use_torch_stream(self):
        """Enqueue on torch's current stream, so kernels are ordered with torch ops and NCCL collectives."""
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        self._chk(self.L.sfe_sync(self.h))

    def launch_count(self):
        return int(self.L.sfe_launch_count(self.h))

    def mask(self):
        m = np.empty(169, np.float32)
        self._chk(self.L.sfe_get_mask(self.h, m.ctypes.data))
        return m

    # ---- pyramids (MakePyramid, hessian.h:95-126 / klt.h:98-137 / brute.h:59-80)
    def pyramid(self, w, h, depth, flavor=HESSIAN, batch=1):
        return Pyramid(self, w, h, depth, flavor, batch)

    def make_pyramid(self, frames, depth, flavor=HESSIAN):
        """frames: (n,H,W,3) or (H,W,3) uint8, numpy (host path) or CUDA torch tensor (device path)."""
        if frames.ndim == 3:
            frames = frames[None]
        n, H, W, _ = frames.shape
        p = Pyramid(self, W, H, depth, flavor, n)
        p.build(frames)
        return p

    # ---- P1 (matcher.cpp:173-206)
    def track_fb(self, pfrom, pto, from_xy, seed_xy, levels=3, thr=0.001, maxit=10, fb_max=0.3, n_per_pair=None,
                 from_first=0, to_first=0, out=None, want_steps=True):
        return self._track(self.L.sfe_track_fb, self.L.sfe_track_fb_dev, pfrom, pto, from_xy, seed_xy, levels, thr,
                           maxit, fb_max, n_per_pair, from_first, to_first, out, want_steps, klt=False)

    # ---- P2 (klt.h:403-424 driven as matcher.cpp:173-206)
    def klt_track_fb(self, pfrom, pto, from_xy, seed_xy, thr=0.001, maxit=10, fb_max=0.3, n_per_pair=None,
                     from_first=0, to_first=0, out=None, want_steps=True):
        return self._track(self.L.sfe_klt_track_fb, self.L.sfe_klt_track_fb_dev, pfrom, pto, from_xy, seed_xy, None,
                           thr, maxit, fb_max, n_per_pair, from_first, to_first, out, want_steps, klt=True)

    def _track(self, fn_host, fn_dev, pfrom, pto, from_xy, seed_xy, levels, thr, maxit, fb_max, n_per_pair,
               from_first, to_first, out, want_steps, klt):
        dev = _is_torch(from_xy) and from_xy.is_cuda
        if dev:
            import torch
            n = from_xy.shape[0]
            to_xy = seed_xy  # updated in place
            if out is None:
                o = dict(back_xy=torch.empty_like(from_xy), status_fwd=torch.empty(n, dtype=torch.int32, device=from_xy.device),
                         status_bwd=torch.empty(n, dtype=torch.int32, device=from_xy.device),
                         accepted=torch.empty(n, dtype=torch.uint8, device=from_xy.device),
                         steps=torch.empty(n, dtype=torch.int32, device=from_xy.device) if want_steps else None)
            else:
                o = out
            lv_arr = levels if (levels is not None and _is_torch(levels)) else None
            fn = fn_dev
        else:
            from_xy = _np(from_xy, np.float32).reshape(-1, 2)
            n = from_xy.shape[0]
            to_xy = _np(seed_xy, np.float32).reshape(-1, 2).copy()
            o = dict(back_xy=np.empty((n, 2), np.float32), status_fwd=np.empty(n, np.int32),
                     status_bwd=np.empty(n, np.int32), accepted=np.empty(n, np.uint8),
                     steps=np.empty(n, np.int32) if want_steps else None)
            lv_arr = None
            if levels is not None and not np.isscalar(levels):
                lv_arr = _np(levels, np.int32)
            fn = fn_host
        npp = n if n_per_pair is None else int(n_per_pair)
        default_levels = int(levels) if (levels is not None and lv_arr is None) else 3
        if klt:
            rc = fn(self.h, pfrom.h, from_first, pto.h, to_first, n, max(npp, 1), _ptr(from_xy), _ptr(to_xy), thr, maxit,
                    fb_max, _ptr(o["back_xy"]), _ptr(o["status_fwd"]), _ptr(o["status_bwd"]), _ptr(o["accepted"]),
                    _ptr(o.get("steps")))
        else:
            rc = fn(self.h, pfrom.h, from_first, pto.h, to_first, n, max(npp, 1), _ptr(from_xy), _ptr(to_xy),
                    _ptr(lv_arr), default_levels, thr, maxit, fb_max, _ptr(o["back_xy"]), _ptr(o["status_fwd"]),
                    _ptr(o["status_bwd"]), _ptr(o["accepted"]), _ptr(o.get("steps")))
        self._chk(rc)
        o = dict(o)
        o["to_xy"] = to_xy
        return o

    def track(self, ptmpl, psearch, tmpl_xy, seed_xy, levels=3, thr=0.001, maxit=10, n_per_pair=None, tmpl_first=0,
              search_first=0):
        """One-directional TrackFeature (hessian.h:243-264 / klt.h:403-424) for a batch (host path)."""
        tmpl_xy = _np(tmpl_xy, np.float32).reshape(-1, 2)
        n = len(tmpl_xy)
        xy = _np(seed_xy, np.float32).reshape(-1, 2).copy()
        lv_arr = None if np.isscalar(levels) else _np(levels, np.int32)
        st = np.empty(n, np.int32)
        steps = np.empty(n, np.int32)
        npp = n if n_per_pair is None else int(n_per_pair)
        self._chk(self.L.sfe_track(self.h, ptmpl.h, tmpl_first, psearch.h, search_first, n, max(npp, 1), _ptr(tmpl_xy),
                                   _ptr(xy), _ptr(lv_arr), int(levels) if lv_arr is None else 3, thr, maxit, _ptr(st),
                                   _ptr(steps)))
        return dict(xy=xy, status=st, steps=steps)

    def get_patches(self, pyr, level, xy, frame=0):
        xy = _np(xy, np.float32).reshape(-1, 2)
        n = len(xy)
        p = np.empty((n, 13, 13), np.float32)
        m = np.empty(n, np.float32)
        q = np.empty(n, np.float32)
        self._chk(self.L.sfe_get_patches(self.h, pyr.h, frame, level, n, _ptr(xy), _ptr(p), _ptr(m), _ptr(q)))
        return p, m, q

    def brute_hessian(self, ptmpl, psearch, level, tmpl_xy, xy, tmpl_frame=0, search_frame=0):
        tmpl_xy = _np(tmpl_xy, np.float32).reshape(-1, 2)
        xy = _np(xy, np.float32).reshape(-1, 2)
        n = len(xy)
        out = np.empty((n, 7), np.float32)
        self._chk(self.L.sfe_brute_hessian(self.h, ptmpl.h, tmpl_frame, psearch.h, search_frame, level, n,
                                           _ptr(tmpl_xy), _ptr(xy), _ptr(out)))
        return out

    def klt_system(self, ptmpl, psearch, level, tmpl_xy, xy, tmpl_frame=0, search_frame=0):
        tmpl_xy = _np(tmpl_xy, np.float32).reshape(-1, 2)
        xy = _np(xy, np.float32).reshape(-1, 2)
        n = len(xy)
        out = np.empty((n, 24), np.float32)
        self._chk(self.L.sfe_klt_system(self.h, ptmpl.h, tmpl_frame, psearch.h, search_frame, level, n,
                                        _ptr(tmpl_xy), _ptr(xy), _ptr(out)))
        return out

    # ---- P3 (brute.h:129-164)
    def brute_track(self, pfrom, pto, from_xy, seed_xy, coarse=BRUTE_COARSE, fine=BRUTE_FINE, n_per_pair=None,
                    from_first=0, to_first=0):
        from_xy = _np(from_xy, np.float32).reshape(-1, 2)
        n = len(from_xy)
        to_xy = _np(seed_xy, np.float32).reshape(-1, 2).copy()
        coarse = _np(coarse, np.float32)
        fine = _np(fine, np.float32)
        st = np.empty(n, np.int32)
        sad = np.empty(n, np.float32)
        pos = C.c_int64(0)
        npp = n if n_per_pair is None else int(n_per_pair)
        self._chk(self.L.sfe_brute_track(self.h, pfrom.h, from_first, pto.h, to_first, n, max(npp, 1), _ptr(from_xy),
                                         _ptr(to_xy), _ptr(coarse), len(coarse) // 2, _ptr(fine), len(fine) // 2,
                                         _ptr(st), _ptr(sad), C.addressof(pos)))
        return dict(to_xy=to_xy, status=st, best_sad=sad, positions=pos.value)

    # ---- batched replay of independent frame pairs (host buffers, chunk-pipelined; include/slamfe.h)
    def replay_pairs(self, frames_from, frames_to, from_xy, seed_xy, depth, levels=3, thr=0.001, maxit=10, fb_max=0.3,
                     n_per_pair=None, chunk_pairs=0, out=None, want_steps=True):
        """frames_*: (npairs,H,W,3) uint8 host arrays (numpy or CPU torch tensors, ideally pinned);
        from_xy/seed_xy: (npairs*n_per_pair, 2) float32 host arrays.  seed_xy is NOT modified."""
        npairs, H, W, ch = frames_from.shape
        assert ch == 3 and tuple(frames_to.shape) == (npairs, H, W, 3)
        from_xy = from_xy if _is_torch(from_xy) else _np(from_xy, np.float32).reshape(-1, 2)
        n = from_xy.shape[0]
        npp = n // max(npairs, 1) if n_per_pair is None else int(n_per_pair)
        assert npp * npairs == n
        if out is None:
            out = dict(to_xy=np.empty((n, 2), np.float32), back_xy=np.empty((n, 2), np.float32), status_fwd=np.empty(n, np.int32),
                       status_bwd=np.empty(n, np.int32), accepted=np.empty(n, np.uint8),
                       steps=np.empty(n, np.int32) if want_steps else None)
        to_xy = out["to_xy"]
        if _is_torch(to_xy):
            to_xy.copy_(seed_xy if _is_torch(seed_xy) else __import__("torch").from_numpy(np.asarray(seed_xy, np.float32)))
        else:
            to_xy[...] = np.asarray(seed_xy, np.float32).reshape(-1, 2)
        lv_arr = None
        if levels is not None and not np.isscalar(levels):
            lv_arr = _np(levels, np.int32)
        for a in (frames_from, frames_to):
            assert (a.is_contiguous() if _is_torch(a) else a.flags.c_contiguous) and not (_is_torch(a) and a.is_cuda)
        self._chk(self.L.sfe_replay_pairs(self.h, W, H, depth, npairs, _ptr(frames_from), _ptr(frames_to), 3 * W, 3 * W * H, max(npp, 1),
                                          _ptr(from_xy), _ptr(to_xy), _ptr(lv_arr), int(levels) if lv_arr is None else 3, thr,
                                          maxit, fb_max, _ptr(out["back_xy"]), _ptr(out["status_fwd"]), _ptr(out["status_bwd"]),
                                          _ptr(out["accepted"]), _ptr(out.get("steps")), int(chunk_pairs)))
        return out

    def replay_sequence(self, frames, pair_stride, from_xy, seed_xy, depth, levels=3, thr=0.001, maxit=10, fb_max=0.3,
                        n_per_pair=None, chunk_pairs=0, out=None, want_steps=True):
        """frames: (nframes,H,W,3) uint8 BGR host array of a replayed sequence -- or (nframes,H,W,2) packed YUYV, the
        camera's native format (sfe_replay_sequence_yuyv); pair i = (frame i, frame i + pair_stride).
        from_xy/seed_xy: ((nframes - pair_stride) * n_per_pair, 2) float32 host arrays, pair-major."""
        nframes, H, W, ch = frames.shape
        npairs = max(nframes - int(pair_stride), 0)
        assert ch in (2, 3) and (frames.is_contiguous() if _is_torch(frames) else frames.flags.c_contiguous)
        from_xy = from_xy if _is_torch(from_xy) else _np(from_xy, np.float32).reshape(-1, 2)
        n = from_xy.shape[0]
        npp = n // max(npairs, 1) if n_per_pair is None else int(n_per_pair)
        assert npp * npairs == n
        if out is None:
            out = dict(to_xy=np.empty((n, 2), np.float32), back_xy=np.empty((n, 2), np.float32), status_fwd=np.empty(n, np.int32),
                       status_bwd=np.empty(n, np.int32), accepted=np.empty(n, np.uint8),
                       steps=np.empty(n, np.int32) if want_steps else None)
        out["to_xy"][...] = np.asarray(seed_xy, np.float32).reshape(-1, 2)
        lv_arr = None
        if levels is not None and not np.isscalar(levels):
            lv_arr = _np(levels, np.int32)
        fn = self.L.sfe_replay_sequence if ch == 3 else self.L.sfe_replay_sequence_yuyv
        self._chk(fn(self.h, W, H, depth, nframes, int(pair_stride), _ptr(frames), ch * W, ch * W * H,
                     max(npp, 1), _ptr(from_xy), _ptr(out["to_xy"]), _ptr(lv_arr),
                     int(levels) if lv_arr is None else 3, thr, maxit, fb_max, _ptr(out["back_xy"]),
                     _ptr(out["status_fwd"]), _ptr(out["status_bwd"]), _ptr(out["accepted"]),
                     _ptr(out.get("steps")), int(chunk_pairs)))
        return out

    # ---- corner seeding (matcher.cpp:313 + :123-130: RGB2GRAY + goodFeaturesToTrack)
    def good_features(self, frames, max_corners=120, quality=0.01, min_distance=20.0, want_eig=False):
        """frames: (n,H,W,3) or (H,W,3) uint8; numpy -> host path (returns per-frame corner arrays, optionally the
        response maps), CUDA tensor -> device path (returns (corners[n,max,2], ncorners[n]) tensors)."""
        if frames.ndim == 3:
            frames = frames[None]
        n, H, W, _ = frames.shape
        if _is_torch(frames) and frames.is_cuda:
            import torch
            xy = torch.zeros((n, max_corners, 2), dtype=torch.float32, device=frames.device)
            cnt = torch.zeros(n, dtype=torch.int32, device=frames.device)
            self._chk(self.L.sfe_good_features_dev(self.h, frames.data_ptr(), W, H, 3 * W, 3 * W * H, n, max_corners,
                                                   float(quality), float(min_distance), xy.data_ptr(), cnt.data_ptr()))
            return xy, cnt
        frames = np.ascontiguousarray(frames, np.uint8)
        xy = np.zeros((n, max_corners, 2), np.float32)
        cnt = np.zeros(n, np.int32)
        eig = np.empty((n, H, W), np.float32) if want_eig else None
        self._chk(self.L.sfe_good_features(self.h, _ptr(frames), W, H, 3 * W, 3 * W * H, n, max_corners, float(quality),
                                           float(min_distance), _ptr(xy), _ptr(cnt), _ptr(eig)))
        corners = [xy[i, :cnt[i]].copy() for i in range(n)]
        return (corners, eig) if want_eig else corners

    # ---- seeding the search (matcher.cpp:224-245) and the live capture format (video.cpp:187-223)
    def seed_features(self, points4, uncertainty, rot4, trans3, k7, from_xy, cols, rows):
        pts = np.ascontiguousarray(points4, np.float64).reshape(-1, 4)
        n = len(pts)
        unc = np.ascontiguousarray(uncertainty, np.float64)
        rot4, trans3, k7 = (np.ascontiguousarray(a, np.float64) for a in (rot4, trans3, k7))
        fxy = _np(from_xy, np.float32).reshape(-1, 2)
        seed, lv, go = np.empty((n, 2), np.float32), np.empty(n, np.int32), np.empty(n, np.uint8)
        self._chk(self.L.sfe_seed_features(self.h, n, _ptr(pts), _ptr(unc), _ptr(rot4), _ptr(trans3), _ptr(k7), _ptr(fxy),
                                           int(cols), int(rows), _ptr(seed), _ptr(lv), _ptr(go)))
        return seed, lv, go

    def yuyv_to_bgr(self, yuyv):
        yuyv = np.ascontiguousarray(yuyv, np.uint8).ravel()
        npx = yuyv.size // 2
        out = np.empty(3 * npx, np.uint8)
        self._chk(self.L.sfe_yuyv_to_bgr(self.h, _ptr(yuyv), npx, _ptr(out)))
        return out

    def pinned(self, shape, dtype):
        """A page-locked host array (sfe_host_alloc): lets the host-pointer entry points copy asynchronously."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._chk(self.L.sfe_host_alloc(self.h, max(nbytes, 1), C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr

    # ---- P4
    # ---- several GPUs of one box (include/slamfe.h; NCCL inside the library, one FrontEnd per rank)
    def dist_init(self, rank, world, id128=None, exchange=None):
        """Creates this rank's NCCL communicator inside the library.  The 128-byte id comes from rank 0
        (dist_unique_id()); `exchange` is a callable that broadcasts a uint8 numpy array from rank 0 (plumbing: e.g. a
        torch.distributed / gloo broadcast) and is used when id128 is not given."""
        if id128 is None:
            id128 = np.zeros(128, np.uint8)
            if rank == 0:
                self._chk(self.L.sfe_dist_unique_id(_ptr(id128)))
            if world > 1:
                id128 = exchange(id128)
        id128 = np.ascontiguousarray(id128, dtype=np.uint8)
        self._chk(self.L.sfe_dist_init(self.h, _ptr(id128), int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def dist_shutdown(self):
        self._chk(self.L.sfe_dist_shutdown(self.h))

    def shard_range(self, n, rank=None, world=None):
        lo, hi = C.c_int64(), C.c_int64()
        rc = self.L.sfe_shard_range(int(n), self.rank if rank is None else rank, self.world if world is None else world,
                                    C.byref(lo), C.byref(hi))
        if rc:
            raise SlamFEError("sfe_shard_range: bad arguments")
        return lo.value, hi.value

    def match_hamming256_sharded(self, q_local, nq_total, t, nt, train_root=0, ratio_num=4, ratio_den=5, max_dist=256, out=None):
        """q_local: this rank's query rows; t: (nt, 8) train rows (valid on train_root, overwritten elsewhere).
        CUDA tensors -> device entry (enqueues on the context's stream), numpy -> host entry.  Returns the gathered
        (idx[nq_total,2], dist[nq_total,2], pass[nq_total])."""
        dev = _is_torch(q_local) and q_local.is_cuda
        if dev:
            import torch
            if out is None:
                out = (torch.empty((nq_total, 2), dtype=torch.int32, device=q_local.device),
                       torch.empty((nq_total, 2), dtype=torch.int32, device=q_local.device),
                       torch.empty(nq_total, dtype=torch.uint8, device=q_local.device))
            fn = self.L.sfe_match_hamming256_sharded_dev
        else:
            q_local = np.ascontiguousarray(q_local).view(np.uint32).reshape(-1, 8)
            t = None if t is None else np.ascontiguousarray(t).view(np.uint32).reshape(-1, 8)
            out = (np.empty((nq_total, 2), np.int32), np.empty((nq_total, 2), np.int32), np.empty(nq_total, np.uint8))
            fn = self.L.sfe_match_hamming256_sharded
        self._chk(fn(self.h, _ptr(q_local), int(nq_total), _ptr(t), int(nt), int(train_root), ratio_num, ratio_den, max_dist,
                     _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    def allgather_rows(self, local, n_total, out=None):
        """local: this rank's rows (CUDA tensor [m, ...]); returns all n_total rows on every rank."""
        import torch
        row = int(np.prod(local.shape[1:])) * local.element_size() if local.ndim > 1 else local.element_size()
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        self._chk(self.L.sfe_allgather_rows_dev(self.h, _ptr(local.contiguous()), row, int(n_total), _ptr(out)))
        return out

    @staticmethod
    def hamming_impl(impl=-1):
        """0 = by size, 1 = integer-ALU kernel, 2 = tensor-core kernel (process-wide); returns the previous setting."""
        return int(lib().sfe_hamming_impl(int(impl)))

    def match_hamming256_async(self, q, t, out, ratio_num=4, ratio_den=5, max_dist=256, batch=1):
        """Host arrays (pinned for real asynchrony); only enqueues -- call sync() before reading `out`."""
        nq, nt = q.shape[0] // batch, t.shape[0] // batch
        self._chk(self.L.sfe_match_hamming256_async(self.h, _ptr(q), nq, _ptr(t), nt, batch, ratio_num, ratio_den, max_dist,
                                                    _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    def match_hamming256(self, q, t, ratio_num=4, ratio_den=5, max_dist=256, batch=1, out=None):
        """q: (batch*nq, 8) uint32, t: (batch*nt, 8) uint32 (numpy -> host path, CUDA tensors -> device path)."""
        dev = _is_torch(q) and q.is_cuda
        if dev:
            import torch
            nq, nt = q.shape[0] // batch, t.shape[0] // batch
            if out is None:
                out = (torch.empty((batch * nq, 2), dtype=torch.int32, device=q.device),
                       torch.empty((batch * nq, 2), dtype=torch.int32, device=q.device),
                       torch.empty(batch * nq, dtype=torch.uint8, device=q.device))
            fn = self.L.sfe_match_hamming256_dev
        else:
            q = np.ascontiguousarray(q).view(np.uint32).reshape(-1, 8)
            t = np.ascontiguousarray(t).view(np.uint32).reshape(-1, 8)
            nq, nt = q.shape[0] // batch, t.shape[0] // batch
            out = (np.empty((batch * nq, 2), np.int32), np.empty((batch * nq, 2), np.int32),
                   np.empty(batch * nq, np.uint8))
            fn = self.L.sfe_match_hamming256
        self._chk(fn(self.h, _ptr(q), nq, _ptr(t), nt, batch, ratio_num, ratio_den, max_dist, _ptr(out[0]), _ptr(out[1]),
                     _ptr(out[2])))
        return out


class Pyramid:
    """sfe_pyr handle: `batch` pyramids of one frame size (Pyramid = vector<GradImage>, hessian.h:42-46)."""

    def __init__(self, fe, w, h, depth, flavor=HESSIAN, batch=1):
        self.fe = fe
        hnd = C.c_void_p()
        fe._chk(fe.L.sfe_pyr_create(fe.h, w, h, depth, flavor, batch, C.byref(hnd)))
        self.h = hnd
        self.w, self.hgt, self.depth, self.flavor, self.batch = w, h, depth, flavor, batch

    def close(self):
        if getattr(self, "h", None) and getattr(self.fe, "h", None):
            self.fe.L.sfe_pyr_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_size(self, level):
        w, h, p = C.c_int(), C.c_int(), C.c_int()
        self.fe._chk(self.fe.L.sfe_pyr_level_size(self.h, level, C.byref(w), C.byref(h), C.byref(p)))
        return w.value, h.value, p.value

    def bytes_per_frame(self):
        return int(self.fe.L.sfe_pyr_bytes_per_frame(self.h))

    def build(self, frames, first=0):
        """MakePyramid for frames (n,H,W,3) uint8: numpy -> sfe_pyr_build (H2D inside), CUDA tensor -> _dev."""
        if frames.ndim == 3:
            frames = frames[None]
        n, H, W, ch = frames.shape
        assert ch == 3 and W == self.w and H == self.hgt
        if _is_torch(frames) and frames.is_cuda:
            assert frames.is_contiguous()
            rc = self.fe.L.sfe_pyr_build_dev(self.fe.h, self.h, frames.data_ptr(), 3 * W, 3 * W * H, first, n)
        else:
            if _is_torch(frames):
                assert frames.is_contiguous()
                ptr = frames.data_ptr()
            else:
                frames = np.ascontiguousarray(frames, np.uint8)
                ptr = frames.ctypes.data
            rc = self.fe.L.sfe_pyr_build(self.fe.h, self.h, ptr, 3 * W, 3 * W * H, first, n)
        self.fe._chk(rc)

    def plane(self, level, frame=0, which=0):
        w, h, _ = self.level_size(level)
        out = np.empty((h, w), np.float32)
        self.fe._chk(self.fe.L.sfe_pyr_download(self.fe.h, self.h, frame, level, which, out.ctypes.data))
        return out
