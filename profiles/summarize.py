"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python profiles/summarize.py launches <launches.csv> <out.md> [title]
  python profiles/summarize.py full <prof.ncu-rep> <out.md> [title]
  python profiles/summarize.py rawcsv <raw.csv> <out.md> [title]      (raw.csv = `ncu -i prof.ncu-rep --page raw --csv`, exported on the GPU box)
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
                "launch__block_size", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum"]


def launches(path, out, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
        f.write(f"{len(data)} launches, {tot / 1e3:.2f} ms total device time.\n\n| share | total us | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * a[1] / tot:.2f}% | {a[1]:.1f} | {a[0]} | `{k[:120]}` |\n")


def full(path, out, title, raw_csv=False):
    txt = open(path).read() if raw_csv else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none` ({path.split('/')[-1]}).\n\n")
        for r in data:
            f.write(f"## `{r[hdr.index('Kernel Name')][:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            if "dram__bytes_read.sum" in hdr:
                pass
            f.write("\n")


if __name__ == "__main__":
    kind, path, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else path
    if kind == "launches":
        launches(path, out, title)
    else:
        full(path, out, title, raw_csv=(kind == "rawcsv"))
