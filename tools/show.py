"""print the key numbers of bench lines / the per-kernel times of an ncu launch list (tools, not product)"""
import collections
import csv
import json
import sys

for f in sys.argv[1:]:
    if f.endswith(".json"):
        d = json.load(open(f))
        r, e = d["roofline"], d.get("e2e")
        print(f, "value %.1fM rows/s" % (d["value"] / 1e6), "ms/step %.3f" % d["ms_per_step"],
              "check %.3f ms idx %.3f ms" % (r["kernel_ms"], r["index_build_ms"]), "frac %.3f" % r["frac"],
              ("e2e %.1fM rows/s %.1f ms h2d %.2f GB" % (e["value"] / 1e6, e["ms_per_step"], e["h2d_bytes_per_step"] / 1e9)) if e else "",
              d["clocks"])
    else:
        rows = [r for r in csv.reader(open(f)) if len(r) > 5]
        hdr, agg = None, collections.OrderedDict()
        for r in rows:
            if r[0] == "ID":
                hdr = r
                continue
            if hdr is None:
                continue
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            agg.setdefault(d["Kernel Name"].split("(")[0], []).append(v)
        for k, v in agg.items():
            print(f"  {k[:60]:60s} n={len(v):3d} last={v[-1] / 1e3:9.1f} us  min={min(v) / 1e3:9.1f} us")
