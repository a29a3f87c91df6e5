"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <round-tag> [full-capture .ncu-rep]

* profiles/<tag>_launches.csv / .md — launch list of `bench.py` under
  `ncu --metrics gpu__time_duration.sum --clock-control none` with per-kernel shares
* profiles/<tag>_<kernel>_ncu.md   — key counters of one `ncu --set full` capture
* profiles/traffic_<tag>.json      — dram bytes per launch of the dominant kernel (bench.py reads it)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")


def launches(tag, csv_name="launches.csv", what="`bench.py --steps 2 --warmup 3`"):
    path = os.path.join(SRC, csv_name)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    with open(os.path.join(OUT, tag + "_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_ns"])
        for r in rows:
            w.writerow([r["ID"], r["Kernel Name"][:90], r["Grid Size"], r["Block Size"], r["Metric Value"].replace(",", "")])
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"][:90], []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, tag + "_launches.md"), "w") as f:
        f.write("# %s — launch list of %s under ncu (gpu__time_duration.sum, --clock-control none)\n\n" % (tag, what))
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total ns | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.0f | %.3f |\n" % (k, len(v), sum(v), sum(v) / tot))


WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # L2 -> SM operand stream (the CTA-pair kernel halves it): bytes the SMs received over the crossbar, of which by TMA loads; L2 sectors and hit rate
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        name = d["Kernel Name"][1].split("(")[0].split("<")[0].split("::")[-1].strip() or "kernel"
        with open(os.path.join(OUT, "%s_%s_ncu.md" % (tag, name)), "w") as f:
            f.write("# %s — `ncu --set full --clock-control none` of `%s`\n\n| metric | unit | value |\n|---|---|---|\n" % (tag, d["Kernel Name"][1][:100]))
            for k in WANT:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")

        def tobytes(x):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[x[0]]
            return float(x[1].replace(",", "")) * scale
        if rd and wr:
            # bench.py reads traffic_<tag>.json for the VQ filter kernel; other kernels get their own file
            fname = "traffic_%s.json" % tag if name.startswith("vq_tc_kernel") else "traffic_%s_%s.json" % (tag, name)
            json.dump({"kernel": name, "dram_bytes_per_launch": tobytes(rd) + tobytes(wr), "read": tobytes(rd), "write": tobytes(wr),
                       "source": os.path.basename(rep)}, open(os.path.join(OUT, fname), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1]
    launches(tag)
    launches(tag + "_pointnet", "launches_pointnet.csv", "`scripts/bench_pointnet.py` (B = 4096, P = 3000)")
    for rep in sys.argv[2:]:
        full(tag, rep)
