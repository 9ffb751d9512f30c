"""Top source/SASS lines by executed instructions from `ncu --page source --csv` output.
usage: python tools/ncu_source_top.py report.ncu-rep [kernel-substring] [topN]"""
import csv, io, subprocess, sys
rep, sub, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else ""), int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
seen = set()
for b in blocks:
    if sub not in b["name"]:
        continue
    hdr = b["rows"][0]
    ie, src, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    rows = [r for r in b["rows"][1:] if len(r) > ie and r[ie].replace(".", "").isdigit()]
    total = sum(float(r[ie]) for r in rows)
    kind = "SASS" if rows and rows[0][0].startswith("0x") or (rows and rows[0][0].isdigit() and len(rows) > 400) else "SRC"
    key = (b["name"], len(rows))
    if key in seen:
        continue
    seen.add(key)
    print(f"## {b['name'][:90]}  [{len(rows)} rows]  total warp-instr = {total:.3e}")
    for r in sorted(rows, key=lambda r: -float(r[ie]))[:top]:
        print(f"  {float(r[ie]) / total * 100:5.1f}%  stall={r[st]:>7s}  {r[src][:110]}")
