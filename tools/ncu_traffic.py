"""Turn an `ncu --metrics ... --csv` log of ONE timed application (bench.py's own command line) into
  profiles/<tag>_app_kernels.md   per-kernel table: launches, time share, DRAM bytes per launch, key percentages
  profiles/traffic.json            per bench stage: measured DRAM bytes per launch (bench.py's roofline.traffic)
A capture that holds more than one application (e.g. every launch of a `--steps 1 --warmup 1` run) is cut down to the
first complete application: from one K-map transform (RowsR2C<..., 0>) to the next.
usage: python tools/ncu_traffic.py gpurun_out/app_metrics_c4.csv r01 "<command that was profiled>" """
import collections
import csv
import json
import os
import re
import sys

src, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= col["Metric Value"]:
        continue
    d = launch.setdefault(int(r[col["ID"]]), {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]]})
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
    d[r[col["Metric Name"]]] = val * scale  # bytes, microseconds, percent


# keep ONE application when the capture holds several
starts = [i for i, d in launch.items() if "RowsR2C<double, 1024, 0>" in d["name"] or "RowsR2C<float, 1024, 0>" in d["name"]]
if len(starts) >= 2:
    launch = collections.OrderedDict((i, d) for i, d in launch.items() if starts[0] <= i < starts[1])
# the sliced contraction runs the same kernels in both directions: the first ozaki_gemm of the application is the
# forward product (preceded by its digit cutting), the second the adjoint one
n_gemm = 0
for d in launch.values():
    if "ozaki_gemm_kernel" in d["name"]:
        d["direction"] = "fwd" if n_gemm == 0 else "adj"
        n_gemm += 1
    elif "ozaki_slice_rows_kernel" in d["name"] or "detector_to_kfast" in d["name"]:
        d["direction"] = "fwd" if n_gemm == 0 else "adj"


def stage_of(name, d=None):
    if d is not None and "direction" in d:
        return "spectral_gemm_" + d["direction"]
    table = [("RowsR2C", "chirpz_rfft"), ("ColsPass<double, 1024, 0,", "chirpz_rfft"), ("ColsPass<float, 1024, 0,", "chirpz_rfft"),
             ("RowsC2R", "chirpz_irfft"), ("ColsPass<double, 1024, 1,", "chirpz_irfft"), ("ColsPass<float, 1024, 1,", "chirpz_irfft"),
             ("lmm_otf_fwd", "lmm_otf_fwd"), ("lmm_otf_adj", "lmm_otf_adj"), ("slit_gather", "slit_gather"),
             ("slit_scatter", "slit_scatter"), ("dgemm_mma_kernel<1, 1>", "spectral_gemm_fwd"),
             ("dgemm_mma_kernel<0, 0>", "spectral_gemm_adj"), ("otgemm_kernel", "spectral_gemm")]
    for key, st in table:
        if key in name:
            return st
    return None


short = lambda n: re.sub(r"^void (surfh::)?", "", re.sub(r"\(.*", "", n))  # noqa: E731
by_kernel = collections.OrderedDict()
for d in launch.values():
    by_kernel.setdefault(short(d["name"]), []).append(d)
total_us = sum(d["gpu__time_duration.sum"] for d in launch.values())
lines = [f"# {tag}: every kernel of one timed application (forward + adjoint), config c4, fp64", "",
         f"Command: `{cmd}` (metrics-only ncu pass; times are cold-cache and serialised: compare SHARES)", "",
         "| kernel | launches | total ms | share | DRAM read / launch | DRAM write / launch | DRAM % | SM % | FP64 pipe % | tensor pipe % | LSU wavefront % | regs | L2 hit % |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
mean = lambda ds, k: sum(d.get(k, 0.0) for d in ds) / len(ds)  # noqa: E731
for k, ds in sorted(by_kernel.items(), key=lambda kv: -sum(d["gpu__time_duration.sum"] for d in kv[1])):
    t = sum(d["gpu__time_duration.sum"] for d in ds)
    lines.append("| `%s` | %d | %.3f | %.3f | %.1f MB | %.1f MB | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.1f |" % (
        k, len(ds), t / 1e3, t / total_us, mean(ds, "dram__bytes_read.sum") / 1e6, mean(ds, "dram__bytes_write.sum") / 1e6,
        mean(ds, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        mean(ds, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        mean(ds, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        mean(ds, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        mean(ds, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        int(mean(ds, "launch__registers_per_thread")), mean(ds, "lts__t_sector_hit_rate.pct")))
lines += ["", f"Total kernel time of the application under ncu: {total_us / 1e3:.2f} ms over {len(launch)} launches."]
out_md = os.path.join(ROOT, "profiles", f"{tag}_app_kernels.md")
open(out_md, "w").write("\n".join(lines) + "\n")

# per-stage traffic: a stage's launch = its kernels run once each (FFT: the cube launches only, i.e. the big ones)
stages = collections.defaultdict(lambda: collections.defaultdict(list))
for d in launch.values():
    st = stage_of(d["name"], d)
    if st is None:
        continue
    stages[st][short(d["name"])].append(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0))
traffic = {}
for st, kernels in stages.items():
    tot, n = 0.0, 0
    for k, vals in kernels.items():
        if st.startswith("chirpz"):
            vals = sorted(vals)[len(vals) // 4:]  # drop the tiny K-map launches
        tot += sum(vals)
        n += len(vals)
    if st.startswith("spectral_gemm"):
        traffic[st] = tot            # the whole stage of one application: digit cutting + the grouped product
    else:
        traffic[st + "_cube" if st.startswith("chirpz") else st] = tot / n  # per kernel launch, like bench.py's stage rows
json.dump({"source": os.path.basename(src), "command": cmd, "unit": "DRAM bytes (read + write) per launch, mean over the captured launches",
           "stages": traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(open(out_md).read())
print(json.dumps(traffic, indent=1))
