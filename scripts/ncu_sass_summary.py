"""Summarise the SASS page of an .ncu-rep: executed warp instructions and stall samples by opcode and by source line.
    python scripts/ncu_sass_summary.py report.ncu-rep [source.cu]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "sass")
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [(int(r[isamp]), int(r[iex]), r[isrc].strip(), r) for r in rows[hi + 1:] if len(r) > iex and r[iex].isdigit()]
    tot, totex = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f"{rows[0][1][:90] if rows and len(rows[0]) > 1 else ''}")
    print(f"warp instructions executed {totex:.4g}, samples {tot}")
    st = collections.Counter()
    for s, e, src, r in data:
        for i, h in stall_cols:
            if r[i].isdigit():
                st[h] += int(r[i])
    print("stalls:", ", ".join(f"{h[6:]} {100 * v / max(1, tot):.1f}%" for h, v in st.most_common(8)))
    op, ops = collections.Counter(), collections.Counter()
    for s, e, src, r in data:
        t = src.split()
        o = t[1] if t[0].startswith("@") else t[0]
        op[o.split(".")[0]] += e
        ops[o.split(".")[0]] += s
    print("opcode        executed     share  samples")
    for o, c in op.most_common(16):
        print(f"{o:12s} {c:12d} {100 * c / totex:6.1f}% {100 * ops[o] / max(1, tot):6.1f}%")
    # by CUDA source line (cuda,sass view: a line row followed by its SASS rows)
    rows = page(rep, "cuda,sass")
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if hi:
        hdr = rows[hi[0]]
        isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        agg = collections.OrderedDict()
        cur = None
        for r in rows[hi[0] + 1:]:
            if len(r) <= iex:
                continue
            if r[0] != "":
                cur = (r[0], r[1].strip()[:100])
                agg.setdefault(cur, [0, 0])
            elif cur is not None and r[iex].isdigit():
                agg[cur][0] += int(r[isamp])
                agg[cur][1] += int(r[iex])
        print("top source lines by executed instructions:")
        for (ln, src), (sm, ex) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"  ex {100 * ex / max(1, totex):5.1f}%  smp {100 * sm / max(1, tot):5.1f}%  L{ln}: {src}")


if __name__ == "__main__":
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    main(sys.argv[1])
