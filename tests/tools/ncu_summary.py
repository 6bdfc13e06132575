"""Summarise an ncu report (--set full) into the JSON kept under profiles/: per-kernel duration, registers, DRAM bytes,
hit rates, issue utilisation and the top stall reasons. Usage: ncu_summary.py report.ncu-rep out.json "description" """
import csv, io, json, subprocess, sys
rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum"]
kern = []
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    k = {"Kernel Name": d["Kernel Name"]}
    for m in keep:
        if m in d and d[m] != "":
            k[m] = f"{d[m]} {u[m]}".strip()
    kern.append(k)
json.dump({"what": what, "kernels": kern}, open(out, "w"), indent=1)
for k in kern:
    print(k["Kernel Name"][:60], k.get("gpu__time_duration.sum"), k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))
