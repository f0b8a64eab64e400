"""Text summary of an ncu --set full report: the raw metrics the roofline discussion uses, per kernel instance, then the
hottest SASS lines of the first instance (tools/ncu_top.py).  Run on the GPU box; the .ncu-rep itself is too big to keep."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'derived__lts__lts2xbar_bytes.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum', 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__inst_executed_op_tma_ld.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-84s %s %s" % (w, r[i], units[i]))
    print()
sys.stdout.flush()
subprocess.run([sys.executable, "tools/ncu_top.py", rep, "12"])
