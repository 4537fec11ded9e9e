#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: executed warp instructions per SASS opcode and stall-sample totals."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
ops = collections.Counter(); stalls = collections.Counter(); total = 0; samples = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr, r))
    src = d["Source"].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(2)
    # collapse modifiers to the family + a few distinguishing suffixes
    fam = op.split(".")[0]
    if fam == "IMAD":
        fam = "IMAD.WIDE" + (".X" if ".X" in op else "") if ".WIDE" in op else ("IMAD.HI" if ".HI" in op else ("IMAD.MOV" if ".MOV" in op else ("IMAD.X" if ".X" in op else "IMAD")))
    elif fam in ("LDL", "STL", "LDS", "STS", "LDG", "STG", "LD", "ST"):
        fam = fam + ("." + op.split(".")[-1] if op.split(".")[-1] in ("128", "64") else "")
    n = int(d["Instructions Executed"] or 0)
    ops[fam] += n; total += n
    samples += int(d["# Samples"] or 0)
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            stalls[k] += int(v)
print("total warp instructions executed: %d ; samples %d" % (total, samples))
for k, v in ops.most_common(30):
    print("  %-16s %12d  %5.1f%%" % (k, v, 100.0 * v / total))
print("stall samples (all):")
for k, v in stalls.most_common(12):
    print("  %-28s %10d  %5.1f%%" % (k, v, 100.0 * v / max(1, samples)))
