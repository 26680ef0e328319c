"""profiles/traffic.json from the ncu CSVs of scratch/gpu_traffic.sh (dram__bytes_read/write.sum, gpu__time_duration.sum per
launch): for every kernel of every configuration the LAST launch (steady state), beside the algorithmic bytes of the phase."""
import csv, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "r02t"
CFG = {"c4": (3, 256, 1), "c3": (3, 203, 3), "c2": (3, 120, 1), "e3": (3, 100, 3), "e2": (2, 2048, 2), "p2d": (2, 2048, 1)}
out = {}
for w, (dim, n, b) in CFG.items():
    path = f"gpurun_out/{tag}_traffic_{w}.csv"
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        launches.setdefault(r[iid], {"kernel": r[ik]})[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    if dim == 3:
        nb_cell, nb_node, nb_edge, nnz = bench.box_counts(n)
        npc = 4
    else:
        nb_node = (n + 1) ** 2; nb_cell = 2 * n * n; nb_edge = 2 * n * (n + 1) + n * n; nnz = nb_node + 2 * nb_edge; npc = 3
    bv, bp = bench.algorithmic_bytes(nb_cell, nb_node, nnz, b=b, npc=npc)
    seen = {}
    for lid in sorted(launches, key=int):
        L = launches[lid]
        name = re.sub(r"^void ", "", L["kernel"])
        name = re.sub(r"afb::", "", name)
        def val(m):
            v, u = L[m]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
            return v * scale
        seen[name] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"), "time_us": val("gpu__time_duration.sum")}
    for name, e in seen.items():
        if not re.match(r"k_(assemble|pattern|scan|row_unique|tile)", name):
            continue
        is_pattern = name.startswith(("k_pattern", "k_scan", "k_row_unique"))
        alg = bp if is_pattern else bv
        e["algorithmic_bytes_of_phase"] = alg
        e["traffic_over_algorithmic"] = (e["dram_bytes_read"] + e["dram_bytes_write"]) / alg
        e["dram_gbs"] = (e["dram_bytes_read"] + e["dram_bytes_write"]) / e["time_us"] / 1e3
        e["source"] = f"profiles/{tag}_traffic_{w}.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; last launch of the kernel; {dim}-D box n={n}, b={b})"
        short = re.sub(r"\(.*$", "", name)
        out[f"{short}:{dim}d:n={n}:b={b}"] = e
# the key bench.py looks up for the bench line of a same-size run (bricks executor)
for k, e in list(out.items()):
    for n in (120, 256):
        if k.startswith(f"k_assemble_tiled<4>:3d:n={n}"):
            out[f"k_assemble_tiled:n={n}"] = e
json.dump(out, open("profiles/traffic.json", "w"), indent=1, sort_keys=True)
for k, e in sorted(out.items()):
    print(f"{k[:70]:70s} {e['time_us']:9.1f} us  R {e['dram_bytes_read']/1e6:8.1f} MB  W {e['dram_bytes_write']/1e6:8.1f} MB  x{e['traffic_over_algorithmic']:.2f} alg  {e['dram_gbs']:7.0f} GB/s")
