"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion needs."""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_uniform', 'lts__t_sector_hit_rate.pct']


def main(path, grep=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    if grep:
        idx = [i for i, h in enumerate(hdr) if grep in h]
    else:
        idx = [i for i, h in enumerate(hdr) if any(h == w or h.startswith(w) for w in WANT)]
    for row in rows[2:]:
        print('---')
        for i in idx:
            print(f'  {hdr[i]} [{units[i]}] = {row[i][:100]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
