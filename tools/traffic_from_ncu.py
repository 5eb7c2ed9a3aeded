"""profiles/r01_final_traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum --csv` log of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e`:
DRAM bytes of the kernels of ONE check (the last classify launch and everything after it = the
check phase; the index phase is listed separately).  usage: traffic_from_ncu.py <csv> <storage> <out.json>"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, per = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    key = (int(d["ID"]), d["Kernel Name"].split("(")[0].replace("void ", "").replace("zk::", ""))
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    per.setdefault(key, {})[d["Metric Name"]] = v * scale
keys = list(per.keys())
INDEX = ("k_set_u32", "k_pos_verify", "k_pos_runlen", "k_slots_clear", "k_index_build", "k_heads_from_offsets")
cls = [i for i, k in enumerate(keys) if k[1].startswith("k_evm_classify")]
# the last COMPLETE check: a classify launch followed by its gate-program kernels and then index kernels again
start = None
for c in reversed(cls):
    end = c + 1
    while end < len(keys) and not keys[end][1].startswith(INDEX) and not keys[end][1].startswith("k_evm_classify"):
        end += 1
    if end < len(keys):
        start = c
        break
check = keys[start:end]
n_check = len(check)
i0 = start
while i0 > 0 and keys[i0 - 1][1].startswith(INDEX):
    i0 -= 1
index = keys[i0:start]
out = {"storage": sys.argv[2], "source": "ncu, " + sys.argv[1], "kernels": {}, "index_kernels": {}}
tot = 0.0
for k in check:
    m = per[k]
    b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    tot += b
    out["kernels"][k[1]] = {"ms": m.get("gpu__time_duration.sum"), "dram_read": m.get("dram__bytes_read.sum"),
                            "dram_write": m.get("dram__bytes_write.sum")}
itot = 0.0
for k in index:
    m = per[k]
    b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    itot += b
    e = out["index_kernels"].setdefault(k[1], {"ms": 0.0, "dram": 0.0, "launches": 0})
    e["ms"] += m.get("gpu__time_duration.sum", 0)
    e["dram"] += b
    e["launches"] += 1
out["dram_bytes_per_check"] = tot
out["dram_bytes_index_phase"] = itot
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps({"dram_bytes_per_check": tot, "dram_bytes_index_phase": itot, "check_kernels": n_check}))
