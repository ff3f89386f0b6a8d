"""Per-source-line view of an .ncu-rep captured with --import-source on (compile with -lineinfo):
    python tools/ncu_lines.py report.ncu-rep [file-substring] [top N]
Prints the lines with the most stall samples / executed instructions and, for bwt.cu-style kernels whose phases are
marked with `// ----` comments, the share of every phase."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
cur, hdr, agg = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0].isdigit() and hdr:
        d = dict(zip(hdr, r))
        num = lambda v: int(v) if v and v.strip("-").isdigit() else 0          # ncu prints "-" for lines without samples
        st = {k[6:]: num(v) for k, v in d.items() if k.startswith("stall_") and "Not" not in k}
        agg.append((cur, int(r[0]), r[1].strip()[:88], num(r[6]), num(r[7]), st))
ts = sum(a[3] for a in agg) or 1
ti = sum(a[4] for a in agg) or 1
print("samples", ts, "warp instructions", ti)
for title, key in (("by samples", 3), ("by instructions", 4)):
    print("---", title)
    for a in sorted(agg, key=lambda a: -a[key])[:top]:
        if only and only not in a[0]:
            continue
        st = " ".join("%s=%d" % kv for kv in sorted(a[5].items(), key=lambda kv: -kv[1])[:3])
        print("%-12s %4d samp %5.1f%% inst %5.1f%% %-44s | %s" % (a[0].split("/")[-1][:12], a[1], 100 * a[3] / ts,
                                                                100 * a[4] / ti, st, a[2]))
# phases
files = {a[0] for a in agg if a[0].endswith(".cu")}
for f in files:
    try:
        src = open(f).read().split("\n")
    except OSError:
        continue
    lines = [a for a in agg if a[0] == f]
    lo, hi = min(a[1] for a in lines), max(a[1] for a in lines)
    marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ----") and lo <= i + 1 <= hi]
    ph, phi = collections.Counter(), collections.Counter()
    for a in lines:
        k = "(before the first mark)"
        for ln, nm in marks:
            if a[1] >= ln:
                k = nm
        ph[k] += a[3]
        phi[k] += a[4]
    print("--- phases of", f.split("/")[-1])
    for k in ph:
        print("%5.1f%% samples %5.1f%% inst  %s" % (100 * ph[k] / ts, 100 * phi[k] / ti, k[:100]))
