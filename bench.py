#!/usr/bin/env python
"""bench.py -- throughput of the front-end hot path on synthetic frames (BASELINE.json metric).

One STEP = one batch of B independent 640x480 frame pairs through the whole path (BASELINE config 2):
  MakePyramid of both frames (hessian.h flavour, 4 levels)            -> pyr_* kernels
  forward/backward patch tracking of 2000 features per pair (P1)      -> track_fb_kernel
  2000 x 2000 256-bit Hamming top-2 + ratio test per pair (P4)        -> hamming_* kernels
`value` is frame pairs/s with all inputs resident in HBM.  `e2e` is the same metric through the host-pointer C ABI
(pinned host buffers, H2D/D2H inside the timed region): the B pairs arrive as a replayed alternating-camera sequence
of B + 2 frames (sfe_replay_sequence: pair i = frames i, i + 2; every frame crosses PCIe and is built once, as
Matcher::Track builds one pyramid per new frame, matcher.cpp:317) plus the descriptor sets of the step.
Before anything is printed, pair 0 of both paths is compared with the CPU oracle (a mismatch is a non-zero exit).

`other_configs` carries the remaining BASELINE configurations, measured in the same run: C1 (one frame, 500 features,
host-pointer ABI: latency), C3 (1920x1080, 5000 features, 8 levels), C4 (this rank's shard of a 65,536-pair replay
through sfe_replay_sequence; a few calls are timed and the shard is projected) and C5 (1M x 1M descriptors, query rows
sharded over the ranks, train set broadcast and result rows all-gathered with NCCL inside libslamfe: strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]
For N > 1 launch with torch.distributed.run (one rank per GPU, weak scaling: every rank processes its own B pairs;
the tracking path has no data-path collective; C5 is the one exchange).
"""
import argparse
import hashlib
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
SEQ_STRIDE = 2
WORKLOAD = "C2: 640x480 frame pairs, 2000 features/pair, 4-level pyramid (both frames) + forward/backward " \
           "Hessian patch tracking + 2000x2000 256-bit Hamming top-2"
METRIC = "tracked frame pairs/sec (pyramid + fwd/bwd track + Hamming match)"
# identical in both arms (the driver compares the dicts): the per-pair workload; batch sizes are reported beside it
CONFIG = {"workload": WORKLOAD, "width": W, "height": H, "features": NFEAT, "levels": LEVELS,
          "l2": "inputs larger than L2 on the GPU arm (hundreds of MB of BGR frames and pyramids per step)"}


def tracker_source_hash():
    """Identifies the build the ncu calibration in profiles/ belongs to: the tracker's sources and build flags."""
    h = hashlib.sha256()
    for f in ("track_hessian.cu", "patch.cuh", "sfe_common.cuh", "build.sh"):
        with open(os.path.join(ROOT, "slam-robot_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def calibration():
    """profiles/traffic.json: per-launch DRAM bytes and warp-instructions per Newton step of the tracking kernel from
    the committed ncu capture, valid only for the build whose source hash it records (tools/ncu_calibrate.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            c = json.load(f)
        c["stale"] = c.get("tracker_source_hash") != tracker_source_hash()
        return c
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peaks():
    """int8 tensor peak for the matcher's roofline: twice the measured dense bf16 rate (Tops/s)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return 2.0 * float(json.load(f)["bf16_tflops"]), "2 x measured bf16 (MEASURED_PEAKS.json)"
    except Exception:
        return 2.0 * 1590.0, "2 x fallback bf16 (B200_PROFILING.md)"


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


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def oracle_check_pair(po, frame_a, frame_b, pts, got, sl, what):
    """One pair of a step against the CPU oracle: positions bit for bit, statuses, accept flags, Newton steps."""
    oa, ob = po.Pyramid(frame_a, DEPTH), po.Pyramid(frame_b, DEPTH)
    o = po.hes_track_fb(oa, ob, pts, pts, LEVELS, THR, MAXIT, FB_MAX)
    bad = [k for k in ("to_xy", "back_xy") if not same_bits(got[k][sl], o[k])]
    bad += [k for k in ("status_fwd", "status_bwd", "accepted") if not np.array_equal(got[k][sl], o[k])]
    if int(np.asarray(got["steps"][sl]).sum()) != o["newton_steps"]:
        bad.append("steps")
    if bad:
        raise SystemExit("bench.py: %s differs from the CPU oracle in %s -- refusing to print a number" % (what, bad))
    return int(o["accepted"].sum())


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
    # one rank per GPU on a multi-socket box: stay on the CPUs next to this rank's GPU, so that the pinned buffers below
    # are first-touched on its NUMA node and the uploads do not cross the socket interconnect (SFE_BENCH_NUMA=0: off).
    # At N = 1 the process keeps all host threads (the cpu_baseline leg uses them).
    numa_cpus = 0
    if world > 1 and os.environ.get("SFE_BENCH_NUMA", "1") != "0":
        numa_cpus = sfe.bind_host_to_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    fe = sfe.FrontEnd(local)
    stream = torch.cuda.Stream(device=dev)
    fe.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        tt = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt]

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

    # ---- parity gate on the measured step itself: pair 0 and pair B-1 of this rank's last step against the CPU oracle
    parity = None
    if rank == 0 and not args.no_check:
        from oracle import pyoracle as po
        host = {k: v.cpu().numpy() for k, v in trk_out.items()}
        host["to_xy"] = to_xy.cpu().numpy()
        hi, hd, hp = (v.cpu().numpy() for v in ham_out)
        checked = []
        for p in sorted({0, B - 1}):
            sl = slice(p * NFEAT, (p + 1) * NFEAT)
            oracle_check_pair(po, frames[p].cpu().numpy(), frames[B + p].cpu().numpy(), pts[sl], host, sl, "resident step, pair %d" % p)
            oi, od, oo = po.hamming256_top2(q[sl], t[sl], *RATIO)
            if not (np.array_equal(hi[sl], oi) and np.array_equal(hd[sl], od) and np.array_equal(hp[sl], oo)):
                raise SystemExit("bench.py: Hamming top-2 of pair %d differs from the CPU oracle" % p)
            checked.append(p)
        parity = {"resident_pairs_checked": checked, "outputs": "pyramid-dependent tracks (positions bit for bit, statuses, accept "
                  "flags, Newton steps) and Hamming idx/dist/pass of those pairs == CPU oracle"}

    # ---- end to end through the host-pointer C ABI (the call a user makes): pinned host inputs, results back in host
    # memory.  The B pairs of a step arrive as a replayed alternating-camera SEQUENCE of B + 2 frames
    # (sfe_replay_sequence: every frame is uploaded and built once); the descriptor matching of the step is enqueued first
    # with sfe_match_hamming256_async and drains with it.
    del frames, pyr
    torch.cuda.empty_cache()
    nfr = B + SEQ_STRIDE
    seq = synth.make_sequence(77 + rank, nfr, H, W, stride=SEQ_STRIDE, device=dev).cpu().pin_memory()
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
    chunk = int(os.environ.get("SFE_BENCH_CHUNK", "0"))

    def step_e2e():
        fe.match_hamming256_async(h_q, h_t, h_ham, *RATIO, batch=B)
        r = fe.replay_sequence(seq, SEQ_STRIDE, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk,
                               chunk_pairs=chunk)
        fe.sync()  # joins the matcher's side stream: every result of the step is in host memory when the step ends
        return r, h_ham

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r, m = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = nfr * H * W * 3 + n * 16 + 2 * n * 32
    d2h = n * (8 + 8 + 4 + 4 + 1 + 4) + n * (8 + 8 + 1)
    if not np.array_equal(m[0], hi if parity else ham_out[0].cpu().numpy()):
        raise SystemExit("bench.py: host and device matching paths disagree")
    if parity is not None:
        from oracle import pyoracle as po
        sq = seq.numpy()
        for p in sorted({0, B - 1}):
            sl = slice(p * NFEAT, (p + 1) * NFEAT)
            oracle_check_pair(po, sq[p], sq[p + SEQ_STRIDE], pts[sl], r, sl, "end-to-end step, pair %d" % p)
        parity["e2e_pairs_checked"] = sorted({0, B - 1})
    e2e_accept = float(np.mean(r["accepted"]))

    ms_total, e2e_ms = max_over_ranks([ms_total, e2e_s * 1e3])
    other = other_configs(args, torch, dist, sfe, synth, fe, dev, rank, world, seq, h_pts, h_trk, max_over_ranks, barrier)

    if rank == 0:
        peak, peak_src = peaks()
        tensor_peak, tensor_src = tensor_peaks()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        pyr_bytes = 2 * B * (3 * W * H + 4 * sum(((W + (1 << l) - 1) >> l) * ((H + (1 << l) - 1) >> l) for l in range(DEPTH)))
        trk_bytes = pyr_bytes - 2 * B * 3 * W * H + n * 45            # read both pyramids once + 45 B/feature
        trk_ms = ms_trk / args.steps
        pyr_ms = ms_pyr / args.steps
        ham_ms = ms_ham / args.steps
        cal = calibration()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        issue_peak = sms * 4 * mhz * 1e6
        # a calibration of another tracker build or batch size is still reported (the instruction count per Newton step
        # moves by a few percent between builds) but flagged: "stale": true means "re-run tools/ncu_calibrate.py"
        usable = cal is not None and "track_fb_kernel" in cal
        stale = usable and (bool(cal["stale"]) or int(cal.get("batch_pairs", -1)) != B)
        per_step = float(cal["track_fb_kernel"]["warp_instructions_per_newton_step"]) if usable else None
        issue_achieved = per_step * newton / (trk_ms * 1e-3) if usable else None
        e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value,
            "unit": "frame pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG, "batch_pairs_per_gpu": B,
            "features_per_sec": world * n * args.steps / (ms_total * 1e-3),
            "matches_per_sec": world * n * args.steps / (ms_total * 1e-3),
            "hamming_comparisons_per_sec_kernel": B * NFEAT * NFEAT / (ham_ms * 1e-3),
            "newton_steps_per_feature": newton / n, "accepted_frac": accepted / n, "ratio_pass_frac": passed / n,
            "phase_ms": {"pyramid": pyr_ms, "track": trk_ms, "hamming": ham_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": "frame pairs/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "entry": "sfe_match_hamming256_async + sfe_replay_sequence (stride %d: %d frames = %d pairs per step) + sfe_sync" % (
                        SEQ_STRIDE, nfr, B),
                    # what each rank's host link actually carried: tells a PCIe limit from a host-DRAM limit as N grows
                    "h2d_gbs_per_rank": h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "accepted_frac": e2e_accept,
                    "host_cpus_bound_rank0": numa_cpus},   # sfe_bind_host_to_device (N > 1): CPUs next to the rank's GPU, 0 = unbound
            # dominant kernel: track_fb_kernel (one launch per step).  It is bound by instruction issue, not by HBM: a
            # pyramid pair is read once (DRAM traffic ~ the algorithmic bytes, <1 % of the HBM peak) and then lives in
            # L1/L2.  achieved = executed warp-instructions per Newton step (ncu capture of THIS build,
            # profiles/traffic.json, refused when the tracker's source hash differs) x Newton steps counted live by the
            # kernel / its CUDA-event time; peak = 4 schedulers x SMs x the SM clock sampled in this run.
            "roofline": {"kernel": "track_fb_kernel<HESSIAN>", "bound": "issue", "achieved": issue_achieved, "peak": issue_peak,
                         "unit": "warp-instructions/s", "frac": issue_achieved / issue_peak if usable else None,
                         "traffic": float(cal["track_fb_kernel"]["dram_bytes"]) if usable else None,
                         "warp_instructions_per_newton_step": per_step,
                         "calibration": None if cal is None else {"file": "profiles/traffic.json", "stale": bool(stale),
                                                                  "batch_pairs": cal.get("batch_pairs"),
                                                                  "tracker_source_hash": cal.get("tracker_source_hash")},
                         "share_of_step": trk_ms / ms_step,
                         "newton_steps_per_sec": newton / (trk_ms * 1e-3),
                         "bilinear_samples_per_sec": (newton * 6 * 169) / (trk_ms * 1e-3),
                         "hbm": {"algorithmic_bytes": trk_bytes, "achieved": trk_bytes / (trk_ms * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": trk_bytes / (trk_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src}},
            # the HBM-streaming kernels of the path
            "roofline_pyramid": {"kernel": "pyr_row_kernel (levels 0+1 fused) + pyr_stream_kernel<down> x2, one build of 2B frames",
                                 "bound": "hbm",
                                 "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / peak,
                                 "traffic": float(cal["pyramid_build"]["dram_bytes"]) if usable and "pyramid_build" in cal else None,
                                 "algorithmic_bytes": pyr_bytes, "peak_source": peak_src,
                                 "share_of_step": pyr_ms / ms_step},
            # the descriptor matcher: an exact int8 contraction on the tensor cores (tcgen05.mma kind::i8, K = 288 per
            # comparison incl. the index block).  achieved = 2 x 288 integer operations per comparison / its CUDA-event
            # time; peak = twice the measured dense bf16 rate (int8 runs at twice the bf16 rate on sm_100).  The ncu capture
            # says what actually bounds it: the ALU pipe of the unpack + top-2 epilogue warps, not the tensor pipe.
            "roofline_hamming": {"kernel": "hamming_mma_kernel (+ hamming_finalize_kernel)", "bound": "tensor",
                                 "achieved": B * NFEAT * NFEAT * 576.0 / (ham_ms * 1e-3) / 1e12, "peak": tensor_peak,
                                 "unit": "Tops/s (int8)", "frac": B * NFEAT * NFEAT * 576.0 / (ham_ms * 1e-3) / 1e12 / tensor_peak,
                                 "peak_source": tensor_src, "comparisons_per_sec": B * NFEAT * NFEAT / (ham_ms * 1e-3),
                                 "traffic": float(cal["hamming_mma_kernel"]["dram_bytes_read"] + cal["hamming_mma_kernel"]["dram_bytes_write"])
                                 if usable and "hamming_mma_kernel" in cal else None,
                                 "limiter": "ALU pipe of the worker warps (profiles/ham_r2*_summary.txt)",
                                 "share_of_step": ham_ms / ms_step},
            "other_configs": other,
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(sample_pairs=args.cpu_pairs)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(args, torch, dist, sfe, synth, fe, dev, rank, world, seq, h_pts, h_trk, max_over_ranks, barrier):
    """BASELINE configs 1, 3, 4, 5 under the same clock (each a short measurement; the details say how short)."""
    if args.no_other:
        return None
    out = {}
    B = args.batch
    # ---- C4: this rank's shard of a 65,536-pair replay through sfe_replay_sequence (host buffers, pipelined)
    total_pairs = 65536
    lo, hi = importlib.import_module("slam-robot_b200.dist").shard_range(total_pairs, rank, world)
    calls = 3
    fe.replay_sequence(seq, SEQ_STRIDE, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(calls):
        fe.replay_sequence(seq, SEQ_STRIDE, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk)
    torch.cuda.synchronize()
    (c4_s,) = max_over_ranks([time.perf_counter() - t0])
    rate = world * B * calls / c4_s
    out["C4_replay_65536_pairs"] = {
        "pairs_per_sec": rate, "unit": "frame pairs/s (pyramid + fwd/bwd track, end to end from pinned host memory)",
        "shard_pairs_per_rank": hi - lo, "timed": "%d calls of sfe_replay_sequence x %d pairs per rank (%.1f %% of the shard), "
        "max over ranks; no data-path collective" % (calls, B, 100.0 * calls * B / max(hi - lo, 1)),
        "projected_seconds_for_65536_pairs": total_pairs / rate}
    # the same replay from the camera's native packed YUYV (video.cpp:187-223 on the device): 2 instead of 3 bytes per
    # pixel cross the host link, which is what limits eight ranks of one box
    yuyv = torch.empty((seq.shape[0], H, W, 2), dtype=torch.uint8).pin_memory()
    yuyv[..., 0] = seq[..., 1]
    yuyv[..., 1] = 128
    fe.replay_sequence(yuyv, SEQ_STRIDE, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(calls):
        fe.replay_sequence(yuyv, SEQ_STRIDE, h_pts, h_pts, DEPTH, LEVELS, THR, MAXIT, FB_MAX, n_per_pair=NFEAT, out=h_trk)
    torch.cuda.synchronize()
    (c4y_s,) = max_over_ranks([time.perf_counter() - t0])
    out["C4_replay_yuyv"] = {"pairs_per_sec": world * B * calls / c4y_s, "unit": "frame pairs/s",
                             "h2d_bytes_per_pair": int(yuyv[0].numel() * yuyv.shape[0] / B),
                             "entry": "sfe_replay_sequence_yuyv (YUYV -> BGR on the device, then the same pipeline)"}
    del yuyv

    # ---- C5: 1M x 1M descriptors, strong scaling: query rows sharded, train broadcast + rows all-gathered by NCCL
    # inside libslamfe (sfe_match_hamming256_sharded_dev); all collectives and kernels on one stream, CUDA-event timed
    nq = nt = args.c5_n
    stream = torch.cuda.Stream(device=dev)
    fe.set_stream(stream.cuda_stream)

    def exchange(id128):
        tns = torch.from_numpy(id128).to(dev)
        dist.broadcast(tns, 0)
        return tns.cpu().numpy()

    fe.dist_init(rank, world, exchange=exchange)
    t_np = synth.make_descriptors(11, nt, dup_frac=0.001)
    q_np = synth.make_descriptors(12, nq, dup_frac=0.01, source=t_np)
    qlo, qhi = fe.shard_range(nq)
    with torch.cuda.stream(stream):
        q_loc = torch.from_numpy(q_np[qlo:qhi].view(np.int32)).to(dev)
        t_root = torch.from_numpy(t_np.view(np.int32)).to(dev) if rank == 0 else None
        t_buf = torch.empty((nt, 8), dtype=torch.int32, device=dev)
        res = (torch.empty((nq, 2), dtype=torch.int32, device=dev), torch.empty((nq, 2), dtype=torch.int32, device=dev),
               torch.empty(nq, dtype=torch.uint8, device=dev))
        ms = []
        for it in range(2):   # the first pass warms NCCL's channels up
            if rank == 0:
                t_buf.copy_(t_root)
            else:
                t_buf.zero_()   # the train set must really arrive through the broadcast
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fe.match_hamming256_sharded(q_loc, nq, t_buf, nt, 0, *RATIO, out=res)
            e1.record(stream)
            barrier()
            ms.append(e0.elapsed_time(e1))
        (c5_ms,) = max_over_ranks([ms[-1]])
        # identical results on every rank: a checksum of the gathered arrays, compared across ranks
        chk = (res[0].to(torch.int64) * 1000003 + res[1].to(torch.int64)).sum() + res[2].to(torch.int64).sum() * 7
        chks = [torch.zeros_like(chk) for _ in range(world)]
        if world > 1:
            dist.all_gather(chks, chk)
        else:
            chks = [chk]
        same = all(int(c) == int(chks[0]) for c in chks)
        if not same:
            raise SystemExit("bench.py: C5 gathered results differ between ranks")
        oracle_rows = 0
        if rank == 0 and not args.no_check:
            from oracle import pyoracle as po
            rows = np.r_[0:64, nq // 2:nq // 2 + 64, nq - 64:nq]   # rows of the first, a middle and the last shard
            oi, od, oo = po.hamming256_top2(q_np[rows], t_np, *RATIO)
            gi, gd, gp = (v[torch.from_numpy(rows).to(dev)].cpu().numpy() for v in res)
            if not (np.array_equal(gi, oi) and np.array_equal(gd, od) and np.array_equal(gp, oo)):
                raise SystemExit("bench.py: C5 rows differ from the CPU oracle")
            oracle_rows = len(rows)
    fe.dist_shutdown()
    fe.set_stream(None)
    out["C5_hamming_1Mx1M"] = {
        "seconds": c5_ms * 1e-3, "comparisons_per_sec": nq * nt / (c5_ms * 1e-3), "nq": nq, "nt": nt, "scaling": "strong",
        "ranks": world, "collectives": "ncclBroadcast(train, %d MB) + sharded match + in-place ncclAllGather(idx, dist, pass: %d MB), "
        "issued by libslamfe on the kernels' stream; CUDA events, max over ranks" % (nt * 32 >> 20, nq * 17 >> 20),
        "results_identical_across_ranks": same, "oracle_rows_checked": oracle_rows}
    del q_loc, t_buf, res, t_root
    torch.cuda.empty_cache()

    if rank != 0:   # C1 and C3 are single-GPU shapes ("replicas only"): rank 0 measures them while the others wait
        barrier()
        return out

    # ---- C1: the live robot's shape -- ONE 640x480 frame, 500 features, host-pointer ABI (frame upload + 6-level
    # pyramid + forward/backward tracking at 3 levels + results back), and the keyframe's corner seeding
    A1, B1 = synth.make_pairs(5, 1, H, W)
    A1, B1 = A1.numpy(), B1.numpy()
    p1 = synth.make_features(9, 500, H, W, margin=16.0).astype(np.float32)
    hA, hB = fe.pinned(A1.shape, np.uint8), fe.pinned(B1.shape, np.uint8)
    hA[...] = A1
    hB[...] = B1
    pa, pb = fe.pyramid(W, H, 6, sfe.HESSIAN, 1), fe.pyramid(W, H, 6, sfe.HESSIAN, 1)
    pa.build(hA)
    c1 = {}
    for levels in (3, 6):
        def frame():
            pb.build(hB)                                   # the new frame's pyramid (host pointer: H2D inside)
            return fe.track_fb(pa, pb, p1, p1.copy(), levels, THR, MAXIT, FB_MAX, n_per_pair=500)
        for _ in range(5):
            frame()
        t0 = time.perf_counter()
        for _ in range(50):
            r1 = frame()
        c1["us_per_frame_%d_levels" % levels] = (time.perf_counter() - t0) / 50 * 1e6
    for _ in range(3):
        fe.good_features(hB[0], 120, 0.01, 20.0)
    t0 = time.perf_counter()
    for _ in range(20):
        corners = fe.good_features(hB[0], 120, 0.01, 20.0)
    c1["us_corner_seeding_120"] = (time.perf_counter() - t0) / 20 * 1e6
    c1.update({"features": 500, "accepted_frac": float(np.mean(r1["accepted"])), "corners": int(len(corners[0])),
               "what": "wall clock per call through the host-pointer C ABI from Python (50 / 20 repetitions), one B200"})
    out["C1_one_frame_500_features"] = c1
    pa.close()
    pb.close()

    # ---- C3: 1920x1080, 5000 features per frame, 8-level pyramid; a step = 16 pairs (32 pyramids + tracking), resident
    W3, H3, NF3, D3, B3 = 1920, 1080, 5000, 8, 16
    stream = torch.cuda.Stream(device=dev)
    fe.set_stream(stream.cuda_stream)
    a3, b3 = synth.make_pairs(100, B3, H3, W3, device=dev)
    fr3 = torch.cat([a3, b3]).contiguous()
    del a3, b3
    pts3 = np.concatenate([synth.make_features(7 + p, NF3, H3, W3, margin=16.0) for p in range(B3)]).astype(np.float32)
    f3 = torch.from_numpy(pts3).to(dev)
    t3 = f3.clone()
    pyr3 = fe.pyramid(W3, H3, D3, sfe.HESSIAN, 2 * B3)
    with torch.cuda.stream(stream):
        def step3():
            pyr3.build(fr3)
            t3.copy_(f3)
            return fe.track_fb(pyr3, pyr3, f3, t3, D3, THR, MAXIT, FB_MAX, n_per_pair=NF3, from_first=0, to_first=B3)
        for _ in range(2):
            r3 = step3()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        pyr3.build(fr3)
        e[1].record(stream)
        for _ in range(3):
            r3 = step3()
        e[2].record(stream)
        torch.cuda.synchronize()
    ms_p3, ms3 = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]) / 3
    peak, _ = peaks()
    out["C3_1080p_5000_features_8_levels"] = {
        "pairs_per_sec": B3 / (ms3 * 1e-3), "features_per_sec": B3 * NF3 / (ms3 * 1e-3), "ms_per_step": ms3, "pairs_per_step": B3,
        "pyramid_ms": ms_p3, "pyramid_hbm_frac": 2 * B3 * pyr3.bytes_per_frame() / (ms_p3 * 1e-3) / 1e9 / peak,
        "newton_steps_per_feature": float(r3["steps"].sum().item()) / (B3 * NF3),
        "accepted_frac": float(r3["accepted"].float().mean().item()), "what": "inputs resident, CUDA events, 3 steps, one B200"}
    fe.set_stream(None)
    pyr3.close()
    barrier()
    return out


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


def single_thread_rate(po, A, B, pts, q, t):
    """What the reference itself does: its front-end is single-threaded."""
    t1 = time.perf_counter()
    cpu_step(po, A[:1], B[:1], pts[:NFEAT], q[:NFEAT], t[:NFEAT], nthreads=1)
    return 1.0 / (time.perf_counter() - t1)


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
    return {"value": sample_pairs / dt, "unit": "frame pairs/s", "cores": cores, "kind": "port",
            "sample": "%d frame pairs of the same workload, OpenMP over features (%d threads); flags -O3 -ffast-math "
                      "-march=native (reference Makefile:4)" % (sample_pairs, cores),
            "single_thread_value": single_thread_rate(po, A, B, pts, q, t)}


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
    sample = "%d frame pairs per step on %d host threads (oracle port, -O3 -ffast-math -march=native; the reference's own " \
             "front-end is single-threaded: single_thread_value)" % (pairs, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "frame pairs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": CONFIG, "batch_pairs_per_step": pairs,
        "cpu_baseline": {"value": value, "unit": "frame pairs/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_thread_value": single_thread_rate(po, A, B, pts, q, t)},
        "e2e": {"value": value, "unit": "frame pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="frame pairs per step per GPU")
    ap.add_argument("--cpu-pairs", type=int, default=4, help="frame pairs in the bounded CPU sample")
    ap.add_argument("--c5-n", type=int, default=1 << 20, help="query and train descriptors of the C5 phase")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block (C1, C3, C4, C5)")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle comparison of the measured step (experiments only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
