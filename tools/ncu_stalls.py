import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
keys = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0][:48]
    vals = sorted(((float(r[h.index(k)]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in keys), reverse=True)[:6]
    print(f"{name:48s} stall cycles per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in vals))
