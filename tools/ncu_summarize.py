#!/usr/bin/env python
"""Turn ncu artefacts (gpurun_out/*.ncu-rep, launches.csv) into the small text summaries kept
under profiles/.   usage:
  python tools/ncu_summarize.py rep  <file.ncu-rep> <out.md> [title]
  python tools/ncu_summarize.py list <launches.csv> <out.md> [title]
"""
import collections, csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def rep(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --set full --clock-control none), one block per captured launch\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
            f.write(f"\n## {d.get('Kernel Name', '?')[:120]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            stalls = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v),
                            reverse=True)[:6]
            f.write("\nTop warp stalls (warps per issue-active cycle): " + ", ".join(
                f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls) + "\n")
            try:
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                tot = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
                f.write(f"\nDRAM traffic per launch: {tot / 1e6:.1f} MB (read + write)\n")
            except Exception:
                pass
    print(open(out).read())


def lst(path, out, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v = v / 1000 if r[ix["Metric Unit"]] in ("ns", "nsecond") else v
        name = r[ix["Kernel Name"]].split("(")[0][:90]
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n\n| share | us | launches | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {100 * v[1] / tot:.1f}% | {v[1]:.1f} | {v[0]} | `{k}` |\n")
        ours = sum(v[1] for k, v in agg.items() if "afr::" in k)
        f.write(f"\nshare of afr:: kernels: {100 * ours / tot:.1f}% of the time, {sum(v[0] for k, v in agg.items() if 'afr::' in k)} launches\n")
        f.write("\nMost-launched kernels:\n\n| launches | us | kernel |\n|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
            f.write(f"| {v[0]} | {v[1]:.1f} | `{k}` |\n")
    print(open(out).read())


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu summary")
