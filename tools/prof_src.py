#!/usr/bin/env python
"""Per-source-line samples / instructions of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
    python tools/prof_src.py rep.ncu-rep <kernel-id> [top-n]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", *(["--kernel-id", f":::{kid}"] if kid.isdigit() else ["--kernel-name", f"regex:{kid}"])],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, H, items = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        H = r; continue
    if H and len(r) == len(H) and r[0] != "":
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(H, r))
        items.append((cur, ln, int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0), r[1][:110], d))
ts = sum(x[2] for x in items) or 1; ti = sum(x[3] for x in items) or 1
print("total samples", ts, "warp instructions", ti)
byfile = {}
for x in items:
    a = byfile.setdefault(x[0], [0, 0]); a[0] += x[2]; a[1] += x[3]
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} samples {v[0]/ts:6.1%} inst {v[1]/ti:6.1%}")
stall_cols = [h for h in (H or []) if h.startswith("stall_")]
for x in sorted(items, key=lambda x: -x[2])[:n]:
    d = x[5]
    top = sorted(((int(d[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{x[0]:18s}:{x[1]:5d} samp={x[2]/ts:6.1%} inst={x[3]/ti:6.1%} {top[0][1][6:]}={top[0][0]} {top[1][1][6:]}={top[1][0]} | {x[4]}")
