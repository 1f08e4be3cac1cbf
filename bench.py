#!/usr/bin/env python
"""bench.py -- throughput of the front-end hot path on synthetic frames (BASELINE.json metric).

One STEP = one batch of B independent 640x480 frame pairs through the whole path:
  MakePyramid of both frames (hessian.h flavour, 4 levels)            -> pyr_* kernels
  forward/backward patch tracking of 2000 features per pair (P1)      -> track_fb_kernel
  2000 x 2000 256-bit Hamming top-2 + ratio test per pair (P4)        -> hamming_* kernels
`value` is frame pairs/s with all inputs resident in HBM; `e2e` is the same metric through the
host-pointer C ABI (pinned host buffers, H2D/D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]
For N > 1 launch with torch.distributed.run (one rank per GPU, weak scaling: every rank
processes its own B pairs; the path has no data-path collective).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, NFEAT, LEVELS, DEPTH = 640, 480, 2000, 4, 4
THR, MAXIT, FB_MAX = 0.001, 10, 0.3
RATIO = (4, 5, 80)
WORKLOAD = "C2: 640x480 frame pairs, 2000 features/pair, 4-level pyramid (both frames) + forward/backward " \
           "Hessian patch tracking + 2000x2000 256-bit Hamming top-2"


def measured_traffic(kernel, batch):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic_r1.json), only when the
    capture was taken at this batch size; None otherwise."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r1.json")) as f:
            e = json.load(f)[kernel]
        return float(e["dram_bytes"]) if int(e["batch_pairs"]) == int(batch) else None
    except Exception:
        return None


def issue_fraction(newton_steps, trk_ms, sms, clocks):
    """Issue-slot utilisation of the tracking kernel: warp-instructions per second against 4 schedulers per SM x SM clock."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r1.json")) as f:
            per_step = float(json.load(f)["track_fb_kernel"]["ncu_warp_instructions_per_newton_step"])
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        achieved = per_step * newton_steps / (trk_ms * 1e-3)
        peak = sms * 4 * mhz * 1e6
        return {"bound": "instruction issue", "achieved": achieved, "peak": peak, "unit": "warp-instructions/s",
                "frac": achieved / peak, "warp_instructions_per_newton_step": per_step,
                "source": "profiles/track_r1i_summary.txt (ncu) x Newton steps counted by the kernel in this run"}
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML from a
    thread every 2 ms (the timed region is ~100 ms; nvidia-smi -lms cannot sample that fast), nvidia-smi as a
    fallback when NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mask, self.power = [], 0, []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.idx
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.idx])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.thread = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.thread.start()

    def _run_nvml(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().split(",")]
                self.sm.append(float(r[0]))
                self.sm_max = float(r[1])
                for bit, v in zip(bits, r[2:6]):
                    if v.lower().startswith("active"):
                        self.mask |= bit
            except Exception:
                time.sleep(0.05)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "source": "nvml" if self.nvml else "nvidia-smi",
                "reasons": reasons}


def make_inputs(torch, synth, batch, device, seed):
    """B frame pairs + features + descriptors, generated on the device in chunks (plumbing)."""
    As, Bs = [], []
    for c0 in range(0, batch, 32):
        n = min(32, batch - c0)
        a, b = synth.make_pairs(seed * 1000 + c0, n, H, W, device=device)
        As.append(a)
        Bs.append(b)
    A = torch.cat(As).contiguous()
    B = torch.cat(Bs).contiguous()
    # SFE_BENCH_MARGIN (experiments only): minimum distance of the features from the image border; 70 keeps every patch
    # of every level inside the image, which is how the cost of the border route was measured (profiles/README.md)
    margin = float(os.environ.get("SFE_BENCH_MARGIN", "16"))
    pts = np.concatenate([synth.make_features(seed * 7919 + p, NFEAT, H, W, margin=margin) for p in range(batch)])
    t = synth.make_descriptors(seed * 31 + 1, batch * NFEAT, dup_frac=0.001)
    q = synth.make_descriptors(seed * 31 + 2, batch * NFEAT, dup_frac=0.2, source=t)
    return A, B, pts.astype(np.float32), q, t


def run_gpu(args):
    import torch
    import torch.distributed as dist
    sfe = importlib.import_module("slam-robot_b200")
    synth = importlib.import_module("slam-robot_b200.synth")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the front-end has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    fe = sfe.FrontEnd(local)
    stream = torch.cuda.Stream(device=dev)
    fe.set_stream(stream.cuda_stream)

    A, Bf, pts, q, t = make_inputs(torch, synth, B, dev, seed=1 + rank)
    n = B * NFEAT
    from_xy = torch.from_numpy(pts).to(dev)
    to_xy = torch.empty_like(from_xy)
    q_d, t_d = torch.from_numpy(q.view(np.int32)).to(dev), torch.from_numpy(t.view(np.int32)).to(dev)
    # one pyramid batch of 2B slots: slots [0, B) hold the first frames of the pairs, [B, 2B) the second frames, so
    # that all 2B pyramids of a step are built by ONE sfe_pyr_build_dev call (the frames sit in one device buffer)
    frames = torch.cat([A, Bf]).contiguous()
    del A, Bf
    pyr = fe.pyramid(W, H, DEPTH, sfe.HESSIAN, 2 * B)
    trk_out = dict(back_xy=torch.empty_like(from_xy), status_fwd=torch.empty(n, dtype=torch.int32, device=dev),
                   status_bwd=torch.empty(n, dtype=torch.int32, device=dev), accepted=torch.empty(n, dtype=torch.uint8, device=dev),
                   steps=torch.empty(n, dtype=torch.int32, device=dev))
    ham_out = (torch.empty((n, 2), dtype=torch.int32, device=dev), torch.empty((n, 2), dtype=torch.int32, device=dev),
               torch.empty(n, dtype=torch.uint8, device=dev))

    def step(ev=None):
        if ev: ev[0].record(stream)
        pyr.build(frames)
        if ev: ev[1].record(stream)
        to_xy.copy_(from_xy)  # seed = from_pt (the uncertainty >= 100 branch, matcher.cpp:225)
        fe.track_fb(pyr, pyr, from_xy, to_xy, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, from_first=0, to_first=B,
                    out=trk_out)
        if ev: ev[2].record(stream)
        fe.match_hamming256(q_d, t_d, *RATIO, batch=B, out=ham_out)
        if ev: ev[3].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # clocks / throttle reasons are sampled from the first warm-up step through the timed region and an
        # untimed tail of identical steps (the timed region alone, ~0.1 s, is shorter than a few NVML queries)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step()
        barrier()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        l0 = fe.launch_count()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        for k in range(args.steps):
            step(evs[k])
        t_end.record(stream)
        barrier()
        launches = fe.launch_count() - l0
        if rank == 0:
            t_tail = time.perf_counter()
            while time.perf_counter() - t_tail < 0.5:
                step()
                torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "warm-up + timed region + 0.5 s untimed tail of identical steps"
        barrier()
    ms_total = t_start.elapsed_time(t_end)
    ms_pyr = sum(e[0].elapsed_time(e[1]) for e in evs)
    ms_trk = sum(e[1].elapsed_time(e[2]) for e in evs)
    ms_ham = sum(e[2].elapsed_time(e[3]) for e in evs)
    newton = int(trk_out["steps"].sum().item())
    accepted = int(trk_out["accepted"].sum().item())
    passed = int(ham_out[2].sum().item())

    # ---- end to end through the host-pointer C ABI (the call a user makes): pinned host inputs, results back
    # in host memory.  sfe_replay_pairs pipelines the step in chunks (upload | pyramids + tracking | download);
    # the descriptor matching of the step is enqueued first with sfe_match_hamming256_async and drains with it.
    hA, hB = frames[:B].cpu().pin_memory(), frames[B:].cpu().pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    fe.set_stream(None)
    h_pts = fe.pinned((n, 2), np.float32)
    h_pts[...] = pts
    h_q, h_t = fe.pinned(q.shape, np.uint32), fe.pinned(t.shape, np.uint32)
    h_q[...] = q
    h_t[...] = t
    h_trk = dict(to_xy=fe.pinned((n, 2), np.float32), back_xy=fe.pinned((n, 2), np.float32), status_fwd=fe.pinned((n,), np.int32),
                 status_bwd=fe.pinned((n,), np.int32), accepted=fe.pinned((n,), np.uint8), steps=fe.pinned((n,), np.int32))
    h_ham = (fe.pinned((n, 2), np.int32), fe.pinned((n, 2), np.int32), fe.pinned((n,), np.uint8))

    def step_e2e():
        fe.match_hamming256_async(h_q, h_t, h_ham, *RATIO, batch=B)
        r = fe.replay_pairs(hA, hB, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk,
                            chunk_pairs=int(os.environ.get("SFE_BENCH_CHUNK", "0")))
        fe.sync()  # joins the matcher's side stream: every result of the step is in host memory when the step ends
        return r, h_ham

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r, m = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = 2 * B * H * W * 3 + n * 16 + 2 * n * 32
    d2h = n * (8 + 8 + 4 + 4 + 1 + 4) + n * (8 + 8 + 1)
    assert np.array_equal(r["accepted"], trk_out["accepted"].cpu().numpy()), "host and device paths disagree"
    assert np.array_equal(m[0], ham_out[0].cpu().numpy()), "host and device matching paths disagree"

    # ---- max over ranks
    tt = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(tt[0]), float(tt[1])

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        pyr_bytes = 2 * B * pyr.bytes_per_frame()                      # per step, both pyramids of every pair
        trk_bytes = pyr_bytes - 2 * B * 3 * W * H + n * 45            # read both pyramids once + 45 B/feature
        trk_ms = ms_trk / args.steps
        pyr_ms = ms_pyr / args.steps
        ham_ms = ms_ham / args.steps
        line = {
            "metric": "tracked frame pairs/sec (pyramid + fwd/bwd track + Hamming match)", "value": value,
            "unit": "frame pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_pairs_per_gpu": B, "width": W, "height": H, "features": NFEAT,
                       "levels": LEVELS, "l2": "inputs larger than L2 (%.0f MB BGR + %.0f MB pyramids per step)" % (
                           2 * B * H * W * 3 / 1e6, 2 * B * 4 * sum((W >> l) * (H >> l) for l in range(DEPTH)) / 1e6)},
            "features_per_sec": world * n * args.steps / (ms_total * 1e-3),
            "matches_per_sec": world * n * args.steps / (ms_total * 1e-3),
            "hamming_comparisons_per_sec_kernel": B * NFEAT * NFEAT / (ham_ms * 1e-3),
            "newton_steps_per_feature": newton / n, "accepted_frac": accepted / n, "ratio_pass_frac": passed / n,
            "phase_ms": {"pyramid": pyr_ms, "track": trk_ms, "hamming": ham_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "frame pairs/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            # dominant kernel: track_fb_kernel (one launch per step). Latency/issue bound, NOT HBM bound
            # (pyramids are read once and then live in L1/L2) -- the fraction is reported as asked.
            "roofline": {"kernel": "track_fb_kernel<HESSIAN>", "bound": "hbm", "achieved": trk_bytes / (trk_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": trk_bytes / (trk_ms * 1e-3) / 1e9 / peak,
                         "traffic": measured_traffic("track_fb_kernel", B), "algorithmic_bytes": trk_bytes,
                         "peak_source": peak_src, "share_of_step": trk_ms / ms_step,
                         "bilinear_samples_per_sec": (newton * 6 * 169) / (trk_ms * 1e-3),
                         # what actually bounds it: instruction issue.  Executed warp-instructions per Newton step come
                         # from the committed ncu capture of this kernel, the Newton steps are counted live by the kernel.
                         "issue": issue_fraction(newton, trk_ms, torch.cuda.get_device_properties(dev).multi_processor_count,
                                                 clocks)},
            # the HBM-streaming kernels of the path
            "roofline_pyramid": {"kernel": "pyr_row_kernel (levels 0+1 fused) + pyr_stream_kernel<down> x2, one build of 2B frames",
                                 "bound": "hbm",
                                 "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / peak, "traffic": measured_traffic("pyramid_build", B),
                                 "algorithmic_bytes": pyr_bytes,
                                 "note": "write-bound: the build writes 1.6x what it reads; write-only HBM traffic reaches 3.9 TB/s and a "
                                         "5 read : 8 write mix 5.5 TB/s on this GPU (tools/hbm_mix_probe.py), against 6.5 TB/s for the copy "
                                         "that defines `peak`",
                                 "share_of_step": pyr_ms / ms_step},
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(sample_pairs=args.cpu_pairs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def host_threads():
    """Host threads available to this process (torchrun exports OMP_NUM_THREADS=1: ask the scheduler instead)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_step(po, A, B, pts, q, t, fast=True, nthreads=0):
    """The same step on the host: the oracle port of the reference's CPU path, all host threads."""
    acc = 0
    nthreads = nthreads or host_threads()
    for p in range(len(A)):
        pa = po.Pyramid(A[p], DEPTH, po.FLAVOR_HESSIAN, fast=fast)
        pb = po.Pyramid(B[p], DEPTH, po.FLAVOR_HESSIAN, fast=fast)
        sl = slice(p * NFEAT, (p + 1) * NFEAT)
        r = po.hes_track_fb(pa, pb, pts[sl], pts[sl], LEVELS, THR, MAXIT, FB_MAX, nthreads=nthreads)
        po.hamming256_top2(q[sl], t[sl], *RATIO, nthreads=nthreads, fast=fast)
        acc += int(r["accepted"].sum())
    return acc


def cpu_inputs(pairs, seed=1):
    import torch
    synth = importlib.import_module("slam-robot_b200.synth")
    A, B, pts, q, t = make_inputs(torch, synth, pairs, "cpu", seed)
    return A.numpy(), B.numpy(), pts, q, t


def cpu_baseline(sample_pairs=4):
    """Oracle port (kind "port": the reference itself cannot be built here) with the
    reference's own compiler flags, on a bounded sample of the same workload."""
    from oracle import pyoracle as po
    po.build(fast=True, native=True)  # -march=native for THIS host
    A, B, pts, q, t = cpu_inputs(sample_pairs)
    cores = host_threads()
    cpu_step(po, A[:1], B[:1], pts[:NFEAT], q[:NFEAT], t[:NFEAT])  # warm-up
    t0 = time.perf_counter()
    cpu_step(po, A, B, pts, q, t)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    cpu_step(po, A[:1], B[:1], pts[:NFEAT], q[:NFEAT], t[:NFEAT], nthreads=1)
    dt1 = time.perf_counter() - t1
    return {"value": sample_pairs / dt, "unit": "frame pairs/s", "cores": cores, "kind": "port",
            "sample": "%d frame pairs of the same workload, OpenMP over features (%d threads); flags -O3 -ffast-math "
                      "-march=native (reference Makefile:4)" % (sample_pairs, cores),
            "single_thread_value": 1.0 / dt1}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    po.build(fast=True, native=True)
    pairs = args.cpu_pairs
    A, B, pts, q, t = cpu_inputs(pairs)
    cores = host_threads()
    for _ in range(args.warmup):
        cpu_step(po, A[:1], B[:1], pts[:NFEAT], q[:NFEAT], t[:NFEAT])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(po, A, B, pts, q, t)
    dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    sample = "%d frame pairs per step on %d host threads (oracle port, -O3 -ffast-math -march=native)" % (pairs, cores)
    print(json.dumps({
        "impl": "reference", "metric": "tracked frame pairs/sec (pyramid + fwd/bwd track + Hamming match)", "value": value,
        "unit": "frame pairs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "batch_pairs_per_step": pairs, "width": W, "height": H,
                                        "features": NFEAT, "levels": LEVELS},
        "cpu_baseline": {"value": value, "unit": "frame pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frame pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frame pairs per step per GPU")
    ap.add_argument("--cpu-pairs", type=int, default=4, help="frame pairs in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
