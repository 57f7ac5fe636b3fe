"""Print the headline counters and warp-stall breakdown of every launch in an .ncu-rep (ncu --set full).
python scripts/ncu_stalls.py <report.ncu-rep>"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "local_load", "local_store"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("====", d.get("Kernel Name", "")[:60], "id", d.get("ID"))
    for k in KEYS:
        for kk, v in d.items():
            if kk == k or (k in ("local_load", "local_store") and k in kk and "sum" in kk):
                print(f"  {kk}: {v}")
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") or ("issue_stalled" in k and "ratio" in k and "not_issued" not in k):
            try:
                st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")))
            except ValueError:
                pass
    for v, k in sorted(set(st), reverse=True)[:10]:
        print(f"  stall {k}: {v:.3f}")
