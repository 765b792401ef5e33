"""Top source lines by stall samples, with the dominant stall reasons, from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_stalls.py file.csv [N]"""
import csv, sys
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, agg = None, None, []
for r in csv.reader(open(sys.argv[1])):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 10 and r[0].isdigit() and r[2] == "-":
        try:
            d = {h: r[i] for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            agg.append((int(r[6]), cur, int(r[0]), r[1].strip()[:90], int(r[7]), {k: int(v) for k, v in d.items() if v not in ("", "-") and int(v) > 0}))
        except ValueError:
            pass
tot = sum(a[0] for a in agg)
print("total samples", tot)
for smp, f, ln, src, inst, d in sorted(agg, reverse=True)[:N]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% %-14s:%-4d inst %5.1fM  %s | %s" % (100.0 * smp / tot, f, ln, inst / 1e6, " ".join("%s=%d" % (k[6:], v) for k, v in top), src[:70]))
