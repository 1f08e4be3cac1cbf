"""Write profiles/traffic.json -- the per-launch figures bench.py's `roofline` block needs from ncu -- from a launch list
captured on the box.  The file records the source hash of the tracker build it belongs to; bench.py refuses a stale one.

On the GPU box (one call):
    python bench.py --steps 2 --warmup 1 --no-cpu --no-other > gpurun_out/cal_bench.json &&
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
        --clock-control none --csv --log-file gpurun_out/cal_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu --no-other --no-check
Here:
    python tools/ncu_calibrate.py gpurun_out/cal_launches.csv gpurun_out/cal_bench.json profiles/launches_r2.csv
"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (tracker_source_hash, workload constants)


def main():
    launches_csv, bench_json, keep_as = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else None)
    line = json.loads([l for l in open(bench_json).read().splitlines() if l.startswith("{")][-1])
    B = int(line["batch_pairs_per_gpu"])
    newton_per_launch = line["newton_steps_per_feature"] * B * bench.NFEAT
    rows = [r for r in csv.reader(l for l in open(launches_csv) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iID = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault((r[iID], r[iK]), {})[r[iM]] = float(r[iV].replace(",", ""))
    def pick(sub, want_batch_max=True):
        c = [m for (i, k), m in per.items() if sub in k]
        if not c:
            return None
        big = max(x["smsp__inst_executed.sum"] for x in c)
        c = [x for x in c if x["smsp__inst_executed.sum"] > 0.7 * big]   # the full-batch launches of the resident step
        n = len(c)
        return {"launches_averaged": n,
                "dram_bytes_read": sum(x["dram__bytes_read.sum"] for x in c) / n,
                "dram_bytes_write": sum(x["dram__bytes_write.sum"] for x in c) / n,
                "ncu_time_us": sum(x["gpu__time_duration.sum"] for x in c) / n / 1e3,
                "warp_instructions": sum(x["smsp__inst_executed.sum"] for x in c) / n}
    out = {"note": "per-launch figures of the resident bench step from `ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,"
                   "smsp__inst_executed.sum` (tools/ncu_calibrate.py); DRAM bytes in bytes, times cold-cache and serialised",
           "tracker_source_hash": bench.tracker_source_hash(), "batch_pairs": B,
           "newton_steps_per_launch": newton_per_launch}
    trk = pick("track_fb_kernel")
    trk["dram_bytes"] = trk["dram_bytes_read"] + trk["dram_bytes_write"]
    trk["warp_instructions_per_newton_step"] = trk["warp_instructions"] / newton_per_launch
    out["track_fb_kernel"] = trk
    # one pyramid build of the resident step = a full-batch pyr_row_kernel launch and the pyr_stream_kernel<down> launches
    # that follow it (levels 2, 3, ...) up to the next tracker launch, in launch order
    order = sorted(per.items(), key=lambda kv: int(kv[0][0]))
    big_row = max(m["smsp__inst_executed.sum"] for (i, k), m in order if "pyr_row_kernel" in k)
    builds, cur = [], None
    for (i, k), m in order:
        if "pyr_row_kernel" in k and m["smsp__inst_executed.sum"] > 0.7 * big_row:
            cur = [("pyr_row_kernel", m)]
            builds.append(cur)
        elif "pyr_stream_kernel" in k and cur is not None:
            cur.append(("pyr_stream_kernel level %d" % (len(cur) + 1), m))
        elif "track_fb_kernel" in k:
            cur = None
    pyr = {"builds_averaged": len(builds), "batch_pairs": B}
    for j, (name, _) in enumerate(builds[0]):
        ms = [b[j][1] for b in builds if len(b) > j]
        pyr[name] = {"dram_bytes_read": sum(x["dram__bytes_read.sum"] for x in ms) / len(ms),
                     "dram_bytes_write": sum(x["dram__bytes_write.sum"] for x in ms) / len(ms),
                     "ncu_time_us": sum(x["gpu__time_duration.sum"] for x in ms) / len(ms) / 1e3}
    pyr["dram_bytes"] = sum(v["dram_bytes_read"] + v["dram_bytes_write"] for v in pyr.values() if isinstance(v, dict))
    out["pyramid_build"] = pyr
    for name in ("hamming_mma_kernel", "hamming_kernel", "hamming_finalize_kernel"):
        p = pick(name)
        if p:
            out[name] = p
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    if keep_as:
        shutil.copyfile(launches_csv, os.path.join(ROOT, keep_as))
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
