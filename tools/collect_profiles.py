"""Copies the evidence of one tools/gpu_final_r02.sh run from gpurun_out/ (scratch) into profiles/ (tracked):
bench lines of both arms, the ncu launch list with a per-kernel summary, the `--set full` capture of the fused kernel
reduced to the metrics the documents quote (and profiles/ncu_traffic_r02.json, which bench.py's roofline.traffic reads),
latency sweep, per-workload and per-fit timings.  Usage: python tools/collect_profiles.py r02j"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
CMD = "python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launch_summary(tag):
    src = os.path.join(OUT, "launches_%s.csv" % tag)
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    k, v = head.index("Kernel Name"), head.index("Metric Value")
    tot = {}
    for r in rows:
        t = tot.setdefault(r[k], [0, 0.0])
        t[0] += 1
        t[1] += float(r[v].replace(",", ""))
    total = sum(t[1] for t in tot.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400: " + CMD,
             "(cold-cache, serialised, every launch boundary a full drain: compare SHARES with the 'kernels' shares of profiles/bench_%s.json)" % tag]
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-70s launches %4d  total %12.1f ns  share %5.1f%%" % (name[:70], n, ns, 100 * ns / total))
    shutil.copy(src, os.path.join(PROF, "launches_%s.csv" % tag))
    open(os.path.join(PROF, "launches_%s_summary.txt" % tag), "w").write("\n".join(lines) + "\n")


def full_capture(tag):
    rep = os.path.join(OUT, "prof_fused3_%s.ncu-rep" % tag)
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    rec = {"Kernel Name": vals[head.index("Kernel Name")]}
    for m in METRICS:
        if m in head:
            i = head.index(m)
            rec[m] = "%s %s" % (vals[i], units[i])
    rec["source"] = "ncu --set full --clock-control none --import-source on -k regex:fused3_model_chi2_kernel -s 4 -c 1, %s (gpurun_out/prof_fused3_%s.ncu-rep)" % (CMD, tag)
    json.dump(rec, open(os.path.join(PROF, "ncu_fused3_%s_summary.json" % tag), "w"), indent=1)

    def num(m):
        i = head.index(m)
        x = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    traffic = {"source": "ncu --set full --clock-control none, %s (gpurun_out/prof_fused3_%s.ncu-rep; summary in profiles/ncu_fused3_%s_summary.json; "
                         "the same kernel with the dense left matrix: profiles/ncu_fused3_dense_r02a_summary.json; first-generation kernel: "
                         "profiles/ncu_fused_gen1_r02_summary.json)" % (CMD, tag, tag),
               "batch": 37888,
               "kernels": {"fused_model_chi2_kernel": {
                   "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                   "duration_us_under_ncu": round(num("gpu__time_duration.sum"), 1),
                   "dmma_inst_pct_of_peak": round(num("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"), 1),
                   "fp64_pipe_pct_of_peak": round(num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
                   "registers": int(num("launch__registers_per_thread")), "grid": int(num("launch__grid_size")),
                   "block": int(num("launch__block_size")), "sass_kernel": rec["Kernel Name"].replace("void ", "").split("(")[0]}}}
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic_r02.json"), "w"), indent=1)


def main():
    tag = sys.argv[1]
    for name in ("bench_%s.json", "bench_ref_%s.json", "model_time_%s.txt", "fit_time_%s.txt", "batch_sweep_%s.json"):
        shutil.copy(os.path.join(OUT, name % tag), os.path.join(PROF, name % tag))
    launch_summary(tag)
    full_capture(tag)


if __name__ == "__main__":
    main()
