"""Hottest CUDA source lines of an .ncu-rep (`ncu -i rep --page source --csv --print-source cuda,sass`): stall samples, executed warp
instructions and the top stall reasons per line, plus the stall totals of the kernel.  usage: ncu_src.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []; hdr = None; fname = None
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":        # source-level rows (SASS rows carry an address)
        out.append((fname, r))
isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[isamp]) for f, r in out); toti = sum(num(r[iinst]) for f, r in out)
print("total samples", tot, " warp instructions", toti)
for f, r in sorted(out, key=lambda fr: -num(fr[1][isamp]))[:top]:
    s = num(r[isamp])
    tp = sorted([(num(r[i]), hdr[i][6:]) for i in stalls], reverse=True)[:3]
    print("%-12s %4s %5.1f%% inst %5.1f%%  %-72s %s" % (f.split('/')[-1], r[0], 100 * s / max(tot, 1), 100 * num(r[iinst]) / max(toti, 1),
                                                    r[1].strip()[:72], " ".join("%s:%d" % (b, a) for a, b in tp if a)))
st = {}
for f, r in out:
    for i in stalls: st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + num(r[i])
print("stall totals:", ", ".join("%s %.1f%%" % (k, 100 * v / max(1, sum(st.values()))) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
