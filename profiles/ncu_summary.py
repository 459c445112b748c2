"""Summarise an .ncu-rep (read here without a GPU): per-kernel key metrics + stall breakdown.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=' * 100)
        print(r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-72s %s %s' % (w, r[i], units[i]))
        print('  -- warp stall reasons (avg warps stalled per issue-active cycle)')
        pre, suf = 'smsp__average_warps_issue_stalled_', '_per_issue_active.ratio'
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(pre) and h.endswith(suf):
                try:
                    st.append((float(r[i]), h[len(pre):-len(suf)]))
                except ValueError:
                    pass
        for v, n in sorted(st, reverse=True)[:8]:
            print('     %-28s %.2f' % (n, v))


if __name__ == '__main__':
    main(sys.argv[1])
