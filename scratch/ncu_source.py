"""Per-CUDA-source-line summary of an .ncu-rep captured with --import-source on (-lineinfo build)."""
import csv, io, subprocess, sys, collections, os
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
cur_file = cur_fn = None
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        cur_fn = r[1][:60]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 40:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    g = lambda name: int(float(r[hdr.index(name) - len(hdr)] or 0))
    key = (cur_fn, cur_file, line)
    a = agg.setdefault(key, [0, 0, 0, 0, r[1].strip()[:100], {}])
    a[0] += g('Instructions Executed'); a[1] += g('# Samples'); a[2] += g('L1 Wavefronts Shared'); a[3] += g('L1 Wavefronts Shared Ideal')
    for name in hdr:
        if name.startswith('stall_') and '(' not in name:
            v = g(name)
            if v:
                a[5][name[6:]] = a[5].get(name[6:], 0) + v
for fn in sorted({k[0] for k in agg}):
    items = [(v, k) for k, v in agg.items() if k[0] == fn]
    tot = sum(v[0] for v, _ in items) or 1; ts = sum(v[1] for v, _ in items) or 1; tw = sum(v[2] for v, _ in items) or 1
    print("==", fn, "inst", tot, "samples", ts, "smem wavefronts", tw)
    for v, k in sorted(items, key=lambda x: -x[0][1])[:top]:
        st = ",".join(f"{n}:{c}" for n, c in sorted(v[5].items(), key=lambda x: -x[1])[:3])
        print(f"{v[1]/ts*100:5.1f}% smp {v[0]/tot*100:5.1f}% inst  wf {v[2]/tw*100:5.1f}% (ideal {v[3]/tw*100:4.1f}%) {k[1]}:{k[2]}: {v[4][:70]}  [{st}]")
