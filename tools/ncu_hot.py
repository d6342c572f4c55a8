"""Top stall-sampled SASS instructions of one kernel in an .ncu-rep (source page):
   python tools/ncu_hot.py report.ncu-rep <kernel regex> [N]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0
seen = set()
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        hdr = rows[k + 1]
        col = {h: i for i, h in enumerate(hdr)}
        body = []
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            if len(rows[k]) == len(hdr):
                body.append(rows[k])
            k += 1
        if name in seen:
            continue
        seen.add(name)
        tot = sum(int(r[col["# Samples"]]) for r in body)
        print(f"## {name}\n total samples {tot}, {len(body)} instructions")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[col[h]]) for r in body) for h in stall_cols}
        print(" stalls:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]]))[:n]
        for i in sorted(idx):
            r = body[i]
            top = sorted(((h[6:], int(r[col[h]])) for h in stall_cols), key=lambda kv: -kv[1])[:2]
            print(f"  [{i:5d}] {int(r[col['# Samples']]):6d}  {r[col['Source']].strip():70s} {top}")
    else:
        k += 1
