"""profiles/r02_launches_bench.md, profiles/r02_hot_kernels.md and profiles/tile_emit_traffic.json from the ncu outputs of
scripts/r2_profile.sh (gpurun_out/<tag>_launches.csv, <tag>_hot.ncu-rep, <tag>_span.ncu-rep).
    python scripts/make_profiles.py r2_prof2"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2_prof2"
OUT = "profiles"

# ------------------------------------------------------------------ launch list of one step
rows = list(csv.reader(open(f"gpurun_out/{tag}_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v
        seq.append((r[ki].split("(")[0].replace("void ", "").replace("symb::", ""), v))
names = [n for n, _ in seq]
cd = [i for i, n in enumerate(names) if n.startswith("class_dedup")]
k = cd[2]
a = k
while a > 0 and not names[a - 1].startswith("tile_fixup"):
    a -= 1
b = k
while not names[b].startswith("tile_fixup"):
    b += 1
step = [x for x in seq[a:b + 1] if not x[0].startswith("at::")]
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    e = agg.setdefault(n, [0, 0.0])
    e[0] += 1
    e[1] += v
md = ["# ncu launch list, one product+cleanup step of `bench.py` (round 2)", "",
      f"Source: `gpurun_out/{tag}_launches.csv`, copied to `profiles/r02_launches_bench.csv` (`SYMMER_BENCH_QUICK=1 ncu --metrics "
      "gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1`; cold-cache and serialised: compare SHARES, "
      f"not absolutes). The third device-resident step of the run: {len(step)} launches of this library, {tot / 1e3:.3f} ms summed "
      "(round 1: 22 launches, 9.205 ms). The emission kernel's share agrees with the live CUDA-event measurement of the bench line "
      "(`roofline.kernel_share_of_step`).", "", "| kernel | launches | us | share |", "|---|---:|---:|---:|"]
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |")
md += ["", "`group_kernel` and the three `*_work_kernel` launches run on an empty candidate list here (no cross term of this workload has "
       "a twin); they are sized by device-side counts, so the host never waits for the candidate count."]
open(f"{OUT}/r02_launches_bench.md", "w").write("\n".join(md) + "\n")
shutil.copy(f"gpurun_out/{tag}_launches.csv", f"{OUT}/r02_launches_bench.csv")


# ------------------------------------------------------------------ ncu --set full captures
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(out)))
    return [dict(zip(rws[0], r)) for r in rws[2:]], dict(zip(rws[0], rws[1]))


keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
md = ["# ncu --set full captures of the hot kernels, round 2", "",
      f"Command: `scripts/r2_profile.sh` (`ncu --set full --clock-control none --import-source on`), reports `gpurun_out/{tag}_hot.ncu-rep` "
      f"(collision-free C5/8 step of `bench.py`) and `gpurun_out/{tag}_span.ncu-rep` (28-generator-span operands, `scripts/probe_class.py`). "
      "Times under ncu are cold-cache and serialised; the bench line has the live numbers. Source-level hot spots of a report: "
      "`python scripts/ncu_sass_summary.py <report>`.", ""]
unit_bytes = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
for rep, title in ((f"gpurun_out/{tag}_hot.ncu-rep", "collision-free product (1.25e8 cross terms, all unique)"),
                   (f"gpurun_out/{tag}_span.ncu-rep", "duplicate-heavy product (1.25e8 cross terms, 1.0e8 unique)")):
    recs, units = raw(rep)
    md += [f"## {title}", ""]
    for d in recs:
        name = d["Kernel Name"].split("(")[0]
        md += [f"### `{name}`", "", "| metric | value | unit |", "|---|---:|---|"]
        for key in keys:
            if d.get(key, "") != "":
                md.append(f"| {key} | {d[key]} | {units.get(key, '')} |")
        md.append("")
        if "tile_emit" in name and "_hot" in rep:
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * unit_bytes[units["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * unit_bytes[units["dram__bytes_write.sum"]]
            json.dump({"kernel": "tile_emit_kernel<4>", "source": f"profiles/r02_hot_kernels.md (gpurun_out/{tag}_hot.ncu-rep)",
                       "rows": 125000000, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_row": (rd + wr) / 125000000},
                      open(f"{OUT}/tile_emit_traffic.json", "w"), indent=1)
            md.append(f"DRAM traffic per emitted row: {(rd + wr) / 125000000:.1f} B against 272 B algorithmic (row + coefficient).\n")
open(f"{OUT}/r02_hot_kernels.md", "w").write("\n".join(md))
print(open(f"{OUT}/r02_launches_bench.md").read())
