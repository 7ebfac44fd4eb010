"""Sum DRAM traffic and time per kernel from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
log: `python tools/ncu_traffic.py <log.csv> [units]` -> per-kernel table + totals (divided by `units`, e.g. images per call)."""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mn = h.index("Metric Name"); mu = h.index("Metric Unit"); idc = h.index("ID")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
agg = collections.defaultdict(lambda: dict(n=set(), rd=0.0, wr=0.0, us=0.0))
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]); name = re.sub(r"^void |ezc::|\(anonymous namespace\)::", "", name)
    v = float(r[mv].replace(",", "")) * SCALE.get(r[mu], 1.0)
    a = agg[name]; a["n"].add(r[idc])
    if r[mn] == "dram__bytes_read.sum": a["rd"] += v
    elif r[mn] == "dram__bytes_write.sum": a["wr"] += v
    elif r[mn] == "gpu__time_duration.sum": a["us"] += v
tot = dict(rd=sum(a["rd"] for a in agg.values()), wr=sum(a["wr"] for a in agg.values()), us=sum(a["us"] for a in agg.values()))
print(f"{'kernel':52s} {'launches':>8s} {'us':>10s} {'read MB':>10s} {'write MB':>10s}")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["rd"] + kv[1]["wr"]))[:25]:
    print(f"{k[:52]:52s} {len(a['n']):8d} {a['us']:10.1f} {a['rd'] / 1e6:10.1f} {a['wr'] / 1e6:10.1f}")
print(f"TOTAL us {tot['us']:.1f}  DRAM read {tot['rd'] / 1e6:.1f} MB  write {tot['wr'] / 1e6:.1f} MB  -> per unit ({units:g}): {(tot['rd'] + tot['wr']) / units / 1e6:.2f} MB")
print(json.dumps({"dram_bytes_total": tot["rd"] + tot["wr"], "dram_bytes_per_unit": (tot["rd"] + tot["wr"]) / units, "units": units,
                  "per_kernel": {k: {"launches": len(a["n"]), "us": round(a["us"], 1), "dram_bytes": a["rd"] + a["wr"]} for k, a in agg.items()}}))
