"""Top source lines by warp-instructions executed from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [N]"""
import csv, sys
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = []
for r in csv.reader(open(sys.argv[1])):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 10 and r[0].isdigit() and r[2] == "-":
        try:
            agg.append((int(r[7]), int(r[8]), cur, int(r[0]), r[1].strip()[:110], int(r[6])))
        except ValueError:
            pass
tot = sum(a[0] for a in agg)
print("total warp-insts", tot, "avg threads %.1f" % (sum(a[1] for a in agg) / max(tot, 1)))
for e, t, f, ln, src, smp in sorted(agg, reverse=True)[:N]:
    print("%5.1f%% thr %4.1f smp %6d  %s:%d  %s" % (100.0 * e / tot, t / max(e, 1), smp, f, ln, src))
