"""Aggregates an `ncu --page source --csv` export by source line (stall samples and executed instructions)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections = []; cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = {'file': r[1], 'rows': []}; sections.append(cur); continue
    if cur is not None: cur['rows'].append(r)
grand = 0; out = []
for s in sections:
    hdr = None; data = []
    for r in s['rows']:
        if r and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(r)
    if not hdr: continue
    il, isrc, isamp, iinst = hdr.index("Line No"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    for r in data:
        try: ln = int(r[il]); sm = int(r[isamp] or 0); ie = int(r[iinst] or 0)
        except Exception: continue
        out.append((s['file'].split('/')[-1], ln, r[isrc], sm, ie)); grand += sm
agg = {}
for f, ln, src, sm, ie in out:
    a = agg.setdefault((f, ln), [src, 0, 0]); a[1] += sm; a[2] += ie
toti = sum(a[2] for a in agg.values())
print("total samples", grand, "total warp instructions", toti)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-16s %5d %6.2f%% inst %5.2f%%  %s" % (f, ln, 100 * a[1] / max(1, grand), 100 * a[2] / max(1, toti), a[0].strip()[:100]))
