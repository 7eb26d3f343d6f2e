"""Print a few headline metrics per kernel from an `ncu --page raw --csv` dump."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "lts__t_sectors_op_write.sum", "smsp__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("----", r[ki][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-75s %-8s %s" % (w, units[i], r[i]))
