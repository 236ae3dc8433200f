"""Turn the ncu artefacts gpurun brought back (gpurun_out/) into the committed text summaries under profiles/.

    python tools/summarize_profile.py <tag> [--workload cfg2] [--precision fp16x3|tf32x3]

Inputs : gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list)
         gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture of the hot kernels)
Outputs: profiles/<tag>_launches.txt, profiles/<tag>_ncu_kernels.txt, profiles/traffic_r02.json (per-launch DRAM
         bytes of the tensor-core GEMM, keyed "<workload>@n<GPUs>[@fp16x3]", read by bench.py for roofline.traffic)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "warp instr / cycle / SM"),
]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0]


def launches(tag, out):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    recs = [(short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3) for r in rows]
    starts = [i for i, r in enumerate(recs) if "prep_rows" in r[0]]
    # one eager step = from a prep launch to the next one; take the last complete one made of single prep launches
    step = None
    for a, b in zip(starts, starts[1:]):
        seg = recs[a:b]
        if sum("gemm3x" in r[0] for r in seg) in (2, 3) and sum("prep_rows" in r[0] for r in seg) == 1 and \
                sum("loss_coeffs" in r[0] for r in seg) == 1:
            step = seg
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; python bench.py --steps 2 --warmup 3 "
                f"--no-cpu-baseline --no-graph  (tag {tag})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES. One step (the L2-flush fill excluded):\n")
        if step:
            step = [r for r in step if not (r[0].startswith("at::") and r[3] > 40.0)]
            tot = sum(r[3] for r in step)
            f.write(f"# total {tot:.1f} us in {len(step)} launches\n")
            agg = {}
            for n, g, b, t in step:
                f.write(f"{t:9.1f} us {100 * t / tot:6.1f}%  grid {g:>14s} block {b:>12s}  {n}\n")
                agg[n] = agg.get(n, 0.0) + t
            f.write("# aggregated\n")
            for n, t in sorted(agg.items(), key=lambda kv: -kv[1]):
                f.write(f"{t:9.1f} us {100 * t / tot:6.1f}%  {n}\n")
        else:
            f.write("# (no complete step found)\n")
            for n, g, b, t in recs:
                f.write(f"{t:9.1f} us  grid {g} block {b}  {n}\n")


def kernels(tag, out, workload):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    gemm_traffic = {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on (tag {tag}); one launch per block below.\n")
        f.write("# Captured under the profiler (serialised, cold L2): use for ratios and traffic, not for timing claims.\n")
        for d in data:
            name = short(d[idx["Kernel Name"]])
            f.write(f"\n== {name}  grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}\n")
            for m, label in METRICS:
                if m in idx:
                    f.write(f"   {label:24s} {d[idx[m]]:>14s} {units[idx[m]]}\n")
            if "gemm3x" in name:
                conv = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                tr = sum(float(d[idx[m]]) * conv[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                gemm_traffic.setdefault(name, []).append(tr)
    if gemm_traffic:
        tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
        cur = {}
        if os.path.exists(tpath):
            with open(tpath) as f:
                cur = json.load(f)
        # one step = one launch of each GEMM kernel (forward, fused backward): average over the kinds
        per_kind = {k: sum(v) / len(v) for k, v in gemm_traffic.items()}
        precision = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "fp16x3"
        key = f"{workload}@n1" + ("" if precision == "tf32x3" else f"@{precision}")
        cur[key] = {"dram_bytes_per_launch": sum(per_kind.values()) / len(per_kind),
                         "per_kernel": per_kind, "note": "ncu --set full, cold L2 (flushed before every replay pass)",
                         "source": f"profiles/{tag}_ncu_kernels.txt"}
        with open(tpath, "w") as f:
            json.dump(cur, f, indent=1)


def loss_kernel(tag, out):
    """Full capture of the fused loss kernel at a config-5 chunk (gpurun_out/prof_loss_cfg5_<tag>.ncu-rep), appended."""
    rep = os.path.join(ROOT, "gpurun_out", f"prof_loss_cfg5_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    extra = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
             ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
             ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
             ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait")]
    with open(out, "a") as f:
        f.write("\n# the fused loss kernel at a config-5 chunk (8192 x 16384), python tools/step_time.py 8192 128 128 256\n")
        for d in data:
            f.write(f"\n== {short(d[idx['Kernel Name']])}  grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}\n")
            for m, label in METRICS + extra:
                if m in idx:
                    f.write(f"   {label:24s} {d[idx[m]]:>14s} {units[idx[m]]}\n")


def main():
    tag = sys.argv[1]
    workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "cfg2"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag, os.path.join(ROOT, "profiles", f"{tag}_launches.txt"))
    kernels(tag, os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.txt"), workload)
    loss_kernel(tag, os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.txt"))


if __name__ == "__main__":
    main()
