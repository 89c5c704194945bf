"""Summarise a gpurun visit's ncu outputs into tracked text files under profiles/.

usage: python tools/summarize_ncu.py <tag>     (reads gpurun_out/, writes profiles/<tag>_*.txt)
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("CARLB_PROF_DIR") or os.path.join(ROOT, "profiles")  # (on the GPU box: a directory under gpurun_out/)

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    # host-path (zero-copy) kernels: bytes the L2 moved to / from system memory over PCIe
    "lts__t_sectors_aperture_sysmem_op_write.sum", "lts__t_sectors_aperture_sysmem_op_read.sum",
    "pcie__write_bytes.sum", "pcie__read_bytes.sum",
]


def launch_list(tag):
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_list.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: python bench.py --steps 20 --warmup 5 --fused-only (the driver's command, fused-rollout leg only; first 600 launches)\n")
        f.write(f"# total kernel time {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v[1] / 1e3:12.1f} us  {v[0]:6d} launches  {100 * v[1] / tot:6.2f}%  avg {v[1] / v[0] / 1e3:9.2f} us  {k}\n")
    print("wrote launch list")


def full(tag, rep):
    p = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(p):
        return
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{tag}_{rep}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; report {rep}.ncu-rep (one row per captured launch)\n")
        for r in rows[2:]:
            f.write(f"kernel: {r[hdr.index('Kernel Name')]}\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"  {w:85s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}\n")
            f.write("\n")
    print("wrote", rep)
    # per-launch DRAM traffic of the captured kernel -> bench.py's roofline.traffic
    import json
    try:
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[rd]) * unit[units[rd]] + float(r[wr]) * unit[units[wr]] for r in rows[2:]]
        tp = os.path.join(PROF, f"{tag}_traffic.json")
        d = json.load(open(tp)) if os.path.exists(tp) else {}
        def avg(name):
            if name not in hdr:
                return None
            xs = [float(r[hdr.index(name)]) for r in rows[2:]]
            return sum(xs) / len(xs)

        d[rep] = {"kernel": rows[2][hdr.index("Kernel Name")], "dram_bytes_per_launch": sum(vals) / len(vals),
                  "launches_captured": len(vals),
                  "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "warp_instructions_per_launch": avg("smsp__inst_executed.sum"),
                  "duration_us_under_ncu": avg("gpu__time_duration.sum")}
        # the bench line printed under the profiler tells how many env-steps one captured launch fused
        logp = os.path.join(OUT, rep.replace("prof_", "ncu_") + ".log")
        if os.path.exists(logp):
            for line in open(logp):
                if line.startswith("{") and "fused_steps_per_launch" in line:
                    try:
                        d[rep]["fused_steps_per_launch"] = json.loads(line)["config"]["fused_steps_per_launch"]
                    except Exception:
                        pass
        json.dump(d, open(tp, "w"), indent=1)
    except Exception as e:  # pragma: no cover
        print("traffic summary failed:", e)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag)
    for rep in sorted(f[:-8] for f in os.listdir(OUT) if f.endswith(".ncu-rep")):
        full(tag, rep)
