"""Summarise an `ncu --csv` launch list (gpu__time_duration + dram bytes) into per-kernel shares and the
average DRAM traffic per tensor-core conv launch (profiles/conv_traffic.json, read by bench.py)."""
import collections, csv, json, sys

path, out_json = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
per = collections.defaultdict(lambda: collections.defaultdict(float))
launch = {}
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})[r["Metric Name"]] = v * scale
tot_t = sum(l.get("gpu__time_duration.sum", 0) for l in launch.values())
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for l in launch.values():
    a = agg[l["name"]]
    a[0] += 1
    a[1] += l.get("gpu__time_duration.sum", 0)
    a[2] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
print(f"{'kernel':58s} {'n':>5s} {'us':>10s} {'share':>7s} {'dram MB':>10s}")
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {n:5d} {t:10.1f} {100*t/tot_t:6.1f}% {b/1e6:10.1f}")
conv = [(n, t, b) for k, (n, t, b) in agg.items()
        if "conv_tile" in k or ("umma_conv" in k and "umma_conv_kernel<1>" not in k and "umma_conv_kernel<2>" not in k)]
if out_json and conv:
    n = sum(c[0] for c in conv); b = sum(c[2] for c in conv); t = sum(c[1] for c in conv)
    json.dump({"source": path, "conv_launches": n, "dram_bytes_per_launch": b / n, "conv_us_total": t,
               "conv_share_of_listed_time": t / tot_t}, open(out_json, "w"), indent=1)
    print("conv launches", n, "avg dram MB/launch", b / n / 1e6, "share", t / tot_t)
