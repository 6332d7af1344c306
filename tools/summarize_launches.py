"""ncu launch list (gpu__time_duration.sum per launch, CSV) -> per-kernel share table.
Times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
    python tools/summarize_launches.py launches.csv [--steady]
--steady: only the launches between the first and the last optimizer update found in the list
(whole steady-state steps; drops data generation and the first-step allocations)."""
import collections
import csv
import re
import sys


def main(path, steady=False):
    rows = []
    with open(path, newline='') as fh:
        rd = csv.reader(l for l in fh if not l.startswith('=='))
        hdr = next(rd)
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rd:
            if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
                continue
            val = float(r[ix['Metric Value']].replace(',', ''))
            unit = r[ix['Metric Unit']]
            us = val / 1000 if unit.startswith('n') else val * 1000 if unit.startswith('m') else val
            rows.append((r[ix['Kernel Name']], us))
    n_steps = None
    if steady:
        opt = [i for i, (n, _) in enumerate(rows) if 'FusedOptimizer' in n]
        first = [i for j, i in enumerate(opt) if j == 0 or i - opt[j - 1] > 20]
        if len(first) >= 2:
            rows = rows[first[0]:first[-1]]
            n_steps = len(first) - 1
    agg = collections.OrderedDict()
    for name, us in rows:
        short = re.sub(r'\(.*', '', name)
        short = re.sub(r'^void ', '', short)[:80]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, {tot / 1000:.2f} ms summed device time (cold-cache, serialised)"
          + (f", {n_steps} steady-state step(s): {len(rows) // n_steps} launches and "
             f"{tot / 1000 / n_steps:.2f} ms per step" if n_steps else ""))
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {n} | {us / 1000:.3f} | {100 * us / tot:.1f}% |")
    own = ('spt::', 'tile::', 'fast::', 'umma::')     # the shortened names of namespace spt
    mine = sum(us for k, (n, us) in agg.items() if k.startswith(own))
    print(f"\nown kernels (spt::*): {100 * mine / tot:.1f}% of device time, "
          f"{sum(n for k, (n, us) in agg.items() if k.startswith(own))} launches")


if __name__ == '__main__':
    main(sys.argv[1], '--steady' in sys.argv[2:])
