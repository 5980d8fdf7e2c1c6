"""Summarise an .ncu-rep (read here, no GPU needed): key counters + SIMD-efficiency histogram + stall mix."""
import csv, collections, subprocess, sys
import re
rep = sys.argv[1]
# pipe utilisation (which of ALU / FMA / LSU / XU is busy) and the L1TEX wavefront counters
EXTRA = re.compile(r"^(sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active|sm__pipe_[a-z0-9_]+_cycles_active\.avg\.pct_of_peak_sustained_active|"
                   r"l1tex__data_pipe_lsu_wavefronts(_mem_shared|_mem_lg)?\.sum|l1tex__lsu_writeback_active\.avg\.pct_of_peak_sustained_elapsed|"
                   r"l1tex__data_bank_(reads|writes)\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_sectors_pipe_lsu_mem_local_op_(ld|st)\.sum|"
                   r"l1tex__t_requests_pipe_lsu_mem_local_op_(ld|st)\.sum|l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate\.pct|"
                   r"l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate\.pct)$")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__cycles_elapsed.avg.per_second",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__cycles_active.avg"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for h, u, v in zip(hdr, units, r):
        if h in want or (EXTRA.search(h) and v not in ("", "0", "0.0")):
            print(f"  {h:72s} {v:>18s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi, end = his[0], (his[1] - 2 if len(his) > 1 else len(rows))  # first kernel of the report
hdr, data = rows[hi], [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
f = lambda r, k: float(r[ix[k]] or 0)
tot = sum(f(r, "Instructions Executed") for r in data)
thr = sum(f(r, "Thread Instructions Executed") for r in data)
print(f"warp-insts {tot:.4g}  avg active threads {thr / tot:.2f}")
h = collections.Counter()
for r in data:
    h[int(f(r, "Avg. Threads Executed") // 4) * 4] += f(r, "Instructions Executed")
print("  active-thread histogram:", ", ".join(f"{k}-{k+3}: {v / tot * 100:.1f}%" for k, v in sorted(h.items())))
st = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
t = {c: sum(f(r, c) for r in data) for c in st}
s = sum(t.values())
print("  stalls:", ", ".join(f"{c[6:]} {v / s * 100:.1f}%" for c, v in sorted(t.items(), key=lambda x: -x[1])[:9]))
