"""Turns the ncu artefacts of a gpurun call (gpurun_out/<TAG>_full.ncu-rep, <TAG>_launches.csv) into the committed
summaries under profiles/: <TAG>_launches.csv, <TAG>_ncu_full_kernels.csv, and (with --publish) issue.json and
roofline_traffic.json, which bench.py copies into its JSON line.   usage: make_profiles.py TAG batch_samples [--publish]"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, batch = sys.argv[1], int(sys.argv[2])
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", os.path.join(G, tag + "_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum.per_second", "dram__bytes.sum.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum.per_second", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed"]
idx = [h.index(w) for w in want if w in h]
with open(os.path.join(P, tag + "_ncu_full_kernels.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([h[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_launches.csv"))
if "--publish" in sys.argv:
    def val(r, name):
        i = h.index(name); v = float(r[i].replace(",", "")); u = units[i]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "ns": 1e-6}.get(u, 1)
    def gbs(r, name):
        i = h.index(name); v = float(r[i].replace(",", ""))
        return v * {"Tbyte/s": 1e3, "Gbyte/s": 1.0, "Mbyte/s": 1e-3, "byte/s": 1e-9}.get(units[i], 1.0)
    def secps(r, name):   # sectors per second, whatever time unit ncu chose
        i = h.index(name); v = float(r[i].replace(",", "")); u = units[i]
        return v * {"ns": 1e9, "us": 1e6, "ms": 1e3, "s": 1.0}.get(u.split("/")[-1], 1.0)
    issue, traffic = {}, {"batch_samples": batch, "capture": tag}
    for r in rows[2:]:
        k = r[h.index("Kernel Name")]
        key = "k_walk" if "k_walk" in k else "k_connect" if "k_connect" in k else "k_shadow" if "k_shadow" in k else None
        if not key: continue
        issue[key] = {"kernel": k.split("(")[0].replace("void bdpt::", ""), "ms": val(r, "gpu__time_duration.sum"),
                      "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "lanes_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                      "long_scoreboard_warps_per_issue": val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                      "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "dram_pct_of_peak": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
                      "dram_gbs": gbs(r, "dram__bytes.sum.per_second"), "l2_gbs": 32e-9 * secps(r, "lts__t_sectors.sum.per_second"), "l2_pct_of_peak": val(r, "lts__t_sectors.sum.pct_of_peak_sustained_elapsed"),
                      "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l1_lsu_data_pipe_pct_of_peak": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), "registers": val(r, "launch__registers_per_thread")}
        traffic[key] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    issue["capture"] = f"profiles/{tag}_ncu_full_kernels.csv (ncu --set full --clock-control none, one launch of {batch} samples each)"
    json.dump(issue, open(os.path.join(P, "issue.json"), "w"), indent=1)
    json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, tag + "_ncu_full_kernels.csv")).read()[:600])
