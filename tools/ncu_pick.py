"""Pick the roofline-relevant metrics out of `ncu --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    print("no data"); sys.exit(0)
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg", "gpc__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "dram__bytes.sum"]
for i, h in enumerate(hdr):
    if h in want or "pipe_tensor" in h or h.startswith("dram__bytes") or "lts__t_sectors_srcunit_tex_op_read.sum" == h:
        print("%-75s %s %s" % (h, vals[i], units[i]))
