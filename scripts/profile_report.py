"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

  python scripts/profile_report.py launches <launches.csv> <out.md> [step_index]
  python scripts/profile_report.py kernel   <report.ncu-rep> <out.md> [--traffic-json profiles/emit_traffic.json --rows N]
"""
import collections
import csv
import json
import subprocess
import sys


def read_launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        name = r[ki].replace('symb::', '').replace('void ', '')
        seq.append((name.split('(')[0][:64], v))
    return seq


def launches(path, out, step_index=4):
    seq = read_launches(path)
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith(('pair_records_kernel', 'pair_keys_kernel'))]
    # a step = from the two sketch/ycount launches before the record kernel up to the last kernel of the emission
    emits = [i for i, (n, _) in enumerate(seq) if n.startswith(('emit_', 'tile_fixup'))]
    k = min(step_index, len(starts) - 1)
    lo, hi = starts[k] - 4, emits[k] + 1
    step = seq[lo:hi]
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    with open(out, 'w') as f:
        f.write(f"# ncu launch list, one product+cleanup step (step #{k} of the run)\n\n")
        f.write(f"Source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and "
                f"serialised: compare SHARES, not absolutes). {len(step)} launches, {tot / 1e3:.3f} ms summed.\n\n")
        f.write("| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |\n")
    print(open(out).read())


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__t_sector_hit_rate.pct']


def kernel(rep, out, traffic_json=None, rows_n=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f"# ncu --set full capture: `{rep}`\n\n")
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')].replace('symb::', '')[:100]
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            vals = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
                    vals[w] = (r[i], units[i])
            stalls = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
            top = sorted(((float(r[i] or 0), hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                          for i in stalls), reverse=True)[:5]
            f.write("\nTop stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in top) + "\n\n")
            if traffic_json and rows_n:
                def to_bytes(v, u):
                    v = float(v.replace(',', ''))
                    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]
                tot = to_bytes(*vals['dram__bytes_read.sum']) + to_bytes(*vals['dram__bytes_write.sum'])
                json.dump({"dram_bytes_per_row": tot / rows_n, "rows": rows_n, "dram_bytes": tot, "source": rep},
                          open(traffic_json, 'w'))
                f.write(f"DRAM traffic {tot / 1e9:.3f} GB for {rows_n} emitted rows = {tot / rows_n:.1f} B/row "
                        f"(algorithmic 272 B/row).\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 4)
    else:
        tj = sys.argv[sys.argv.index('--traffic-json') + 1] if '--traffic-json' in sys.argv else None
        rn = int(sys.argv[sys.argv.index('--rows') + 1]) if '--rows' in sys.argv else None
        kernel(sys.argv[2], sys.argv[3], tj, rn)
