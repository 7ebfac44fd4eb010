"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mn = h.index("Metric Name"); mu = h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else v * 1e3 if r[mu] in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"^void |ezc::|\(anonymous namespace\)::", "", name)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'us':>10s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k[:58]:58s} {v[0]:8d} {v[1]:10.1f} {100 * v[1] / tot:5.1f}%")
print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")
