#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: top stall sites per kernel section.

usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv [kernel-substr] [top]
"""
import csv, sys

def sections(path):
    cur = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if row and row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                yield cur
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)

def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    seen = set()
    for sec in list(sections(path)):
        if want not in sec["name"] or not sec["hdr"]:
            continue
        hdr = sec["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
        if "Source" not in idx or "# Samples" not in idx:
            continue
        is_sass = any(r[idx["Source"]].lstrip().split(" ")[0].isupper() for r in sec["rows"][:5] if r[idx["Source"]].strip())
        key = (sec["name"], is_sass)
        if key in seen:
            continue
        seen.add(key)
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        rows = []
        for r in sec["rows"]:
            try:
                s = int(r[idx["# Samples"]])
            except ValueError:
                continue
            rows.append((s, r))
        tot = sum(s for s, _ in rows) or 1
        print(f"=== {sec['name'][:90]} [{'SASS' if is_sass else 'source'}] samples={tot}")
        agg = {c: sum(int(r[idx[c]] or 0) for _, r in rows) for c in stall_cols}
        print("  stall mix:", ", ".join(f"{c[6:]}={100*v/tot:.0f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for s, r in sorted(rows, key=lambda x: -x[0])[:top]:
            reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            print(f"  {100*s/tot:5.1f}%  {r[idx['Source']].strip()[:100]}   [{', '.join(f'{n}:{c}' for c, n in reasons if c)}]  exec={r[idx['Instructions Executed']]}")

if __name__ == "__main__":
    main()
