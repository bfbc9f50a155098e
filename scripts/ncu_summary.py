"""Summarise one .ncu-rep (first profiled launch) into a small CSV of the metrics we track."""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'launch__block_size', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']


def main():
  rep, out = sys.argv[1], sys.argv[2]
  txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(txt.splitlines()))
  hdr, units = rows[0], rows[1]
  # several launches in the report: summarise the longest one
  dur = hdr.index('gpu__time_duration.sum')
  tosec = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}
  vals = max(rows[2:], key=lambda r: float(r[dur].replace(',', '')) if len(r) > dur and r[dur] else 0.0)
  del tosec
  name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
  with open(out, 'w') as f:
    f.write('# %s\n# kernel: %s\nmetric,unit,value\n' % (rep, name))
    for h, u, v in zip(hdr, units, vals):
      if h in KEEP:
        f.write('%s,%s,%s\n' % (h, u, v))
  print(open(out).read())


if __name__ == '__main__':
  main()
