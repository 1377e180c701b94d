"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_x_launches.md "<command>"
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_x_full.md
"""
import collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__pipe_tensor_op_hmma_cycles_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Command: `%s`\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n" % cmd)
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:110], v[0], v[1] / 1e6, 100 * v[1] / tot))
        f.write("\nTotal: %d launches, %.3f ms\n" % (sum(v[0] for v in agg.values()), tot / 1e6))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % src)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## %s  (id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (d.get("Kernel Name", "?")[:120], d.get("ID")))
            for k in hdr:
                if any(k.startswith(p) for p in KEYS):
                    f.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
