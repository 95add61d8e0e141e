#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ directory (scripts/gpu_round.sh) into the tracked evidence under profiles/:
  profiles/<tag>_launches.csv     the ncu launch list of `bench.py --steps 2 --warmup 1`
  profiles/<tag>_launches.md      per-kernel totals / shares of that list
  profiles/<tag>_<kernel>.md      key metrics of the `ncu --set full` capture(s)
  profiles/<tag>_bench.json       the bench lines of the same call (not under a profiler)
usage: python scripts/summarize_profile.py gpurun_out/r01a [prof_name ...]
"""
import collections, csv, json, os, shutil, subprocess, sys

src = sys.argv[1].rstrip("/")
tag = os.path.basename(src)
os.makedirs("profiles", exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def launches():
    p = os.path.join(src, "launches.csv")
    if not os.path.exists(p):
        return
    shutil.copy(p, f"profiles/{tag}_launches.csv")
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"profiles/{tag}_launches.md", "w") as f:
        f.write(f"# {tag}: ncu launch list of `bench.py --steps 2 --warmup 1 --no-cpu-baseline` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Times under the profiler are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {t:.3f} | {t / n:.4f} | {100 * t / tot:.1f}% |\n")
    print(open(f"profiles/{tag}_launches.md").read())


def full(prof):
    rep = os.path.join(src, prof + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(f"profiles/{tag}_{prof}.md", "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` — {prof}\n\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            u = dict(zip(hdr, units))
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            stalls = []
            for k in hdr:
                if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                    try:
                        stalls.append((float(d[k]), k))
                    except ValueError:
                        pass
            f.write("\nWarp stall reasons (warps stalled per issue-active cycle, top 8):\n\n| reason | ratio |\n|---|---:|\n")
            for v, k in sorted(stalls, reverse=True)[:8]:
                f.write(f"| {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |\n")
            try:
                tr = float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
                tw = float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
                f.write(f"\nDRAM traffic per launch: {tr + tw:.4g} B (read {tr:.4g} + write {tw:.4g})\n\n")
                if "k_sims_device" in d.get("Kernel Name", ""):
                    # bench.py's roofline.traffic: the capture is of `bench.py --steps 1 --warmup 3` (4096 games x 800 sims per launch)
                    json.dump({"source": f"profiles/{tag}_{prof}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                               "kernel": d["Kernel Name"], "dram_bytes_per_launch": tr + tw, "dram_bytes_read": tr, "dram_bytes_write": tw,
                               "sims_per_launch": 4096 * 800}, open("profiles/traffic.json", "w"), indent=1)
            except (KeyError, ValueError):
                pass
    print(open(f"profiles/{tag}_{prof}.md").read())


launches()
for prof in sys.argv[2:] or ["prof_sims"]:
    full(prof)
for name in ("bench.json", "bench_reference.json"):
    p = os.path.join(src, name)
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copy(p, f"profiles/{tag}_{name}")
