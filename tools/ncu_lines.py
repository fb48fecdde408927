#!/usr/bin/env python
"""Per-source-line stall-sample / instruction summary of an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [top=40] [kernel-index=0]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = {}; kernel = None; kcount = 0
want_k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        if r[1] != kernel: kernel = r[1]
        continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":      # a source-line summary row
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur, int(r[0]))
        e = out.setdefault(key, {"src": r[1].strip()[:100], "samples": 0, "inst": 0, "stalls": {}})
        e["samples"] += int(d["# Samples"]); e["inst"] += int(d["Instructions Executed"])
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v):
                e["stalls"][k[6:]] = e["stalls"].get(k[6:], 0) + int(v)
tot = sum(e["samples"] for e in out.values()); toti = sum(e["inst"] for e in out.values())
print(f"total samples {tot}, warp-instructions {toti}")
for (f, ln), e in sorted(out.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3])
    print(f"{e['samples']:7d} {100*e['samples']/max(tot,1):5.1f}% inst {100*e['inst']/max(toti,1):5.1f}%  {f}:{ln:<4d} {e['src'][:70]:70s} [{st}]")
