"""Markdown table of the counters this project quotes, from an ncu report's raw page:
    ncu -i report.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_table.py raw.csv [edges]
`edges` (optional): the launch's edge count, adds warp instructions per edge."""
import csv
import sys

METRICS = [
    ('duration under ncu, us', 'gpu__time_duration.sum', 1.0),
    ('warp instructions, M', 'smsp__inst_executed.sum', 1e-6),
    ('issue slots busy, %', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1.0),
    ('resident warps, % of 64', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1.0),
    ('registers / thread', 'launch__registers_per_thread', 1.0),
    ('DRAM read, MB', 'dram__bytes_read.sum', 1.0),
    ('DRAM write, MB', 'dram__bytes_write.sum', 1.0),
    ('DRAM throughput, % of peak', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('L2 hit rate, %', 'lts__t_sector_hit_rate.pct', 1.0),
    ('L2 throughput, % of peak', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('L1/TEX throughput, % of peak', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 1.0),
    ('tensor pipe active, %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1.0),
    ('stall long scoreboard (warps / issue)',
     'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 1.0),
    ('stall short scoreboard', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 1.0),
    ('stall wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 1.0),
    ('stall barrier / mbarrier', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 1.0),
]


def main(path, edges=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    names = []
    for r in kernels:
        n = r[ix['Kernel Name']]
        n = n.split('(')[0].replace('void ', '').split('::')[-1]
        names.append(n)
    print('| metric | ' + ' | '.join(f'`{n}`' for n in names) + ' |')
    print('|---|' + '---:|' * len(names))
    for label, key, scale in METRICS:
        if key not in ix:
            continue
        vals = []
        for r in kernels:
            try:
                v = float(r[ix[key]].replace(',', '')) * scale
                u = units[ix[key]]
                if key.startswith('dram__bytes') and u.lower().startswith('gbyte'):
                    v *= 1000.0
                vals.append(f'{v:.1f}' if v >= 10 else f'{v:.2f}')
            except ValueError:
                vals.append(r[ix[key]])
        print(f'| {label} | ' + ' | '.join(vals) + ' |')
    if edges:
        vals = [f"{float(r[ix['smsp__inst_executed.sum']].replace(',', '')) / edges:.1f}"
                for r in kernels]
        print('| warp instructions per edge | ' + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
