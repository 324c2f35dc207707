"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page).
usage: python profiles/ncu_hot.py rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[isamp]) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[isamp])*100.0/tot:5.1f}% ex={r[iex]:>10s} {r[isrc].strip()[:70]:70s} {st}")
