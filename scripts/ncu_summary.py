"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.txt
  python scripts/ncu_summary.py full gpurun_out/conv_full_r01.ncu-rep profiles/r01_conv_full_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'dram__bytes_write.sum.per_second', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg',
]


def to_us(value, unit):
    value = float(value.replace(',', ''))
    return {'ns': value / 1e3, 'us': value, 'ms': value * 1e3, 's': value * 1e6}.get(unit, value)


def launches(path, out, limit=None):
    with open(path) as handle:
        lines = [line for line in handle if not line.startswith('==')]
    rows = [row for row in csv.DictReader(lines) if row['Metric Name'] == 'gpu__time_duration.sum']
    if limit:
        rows = rows[:int(limit)]  # one step's worth of launches
    total, count = collections.OrderedDict(), collections.Counter()
    for row in rows:
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('unnamed>::', '')
        total[name] = total.get(name, 0.0) + to_us(row['Metric Value'], row['Metric Unit'])
        count[name] += 1
    grand = sum(total.values())
    with open(out, 'w') as handle:
        handle.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none: {len(rows)} launches of one bench '
                     f'step ({path}); serialised cold-cache times: compare SHARES\n')
        handle.write(f'{"kernel":60s} {"launches":>8s} {"total_us":>12s} {"share":>7s}\n')
        for name, value in sorted(total.items(), key=lambda kv: -kv[1]):
            handle.write(f'{name[:60]:60s} {count[name]:8d} {value:12.1f} {100 * value / grand:6.1f}%\n')
        handle.write(f'{"TOTAL":60s} {len(rows):8d} {grand:12.1f}\n')
    print(open(out).read())


def full(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    index = {name: i for i, name in enumerate(header)}
    with open(out, 'w') as handle:
        handle.write(f'# ncu --set full --clock-control none ({path}); one block per captured launch\n')
        for row in rows[2:]:
            handle.write(f'\n{row[index["Kernel Name"]]}\n')
            for metric in FULL_METRICS:
                if metric in index:
                    handle.write(f'  {metric:70s} {row[index[metric]]:>16s} {units[index[metric]]}\n')
    print(open(out).read())




def traffic(path, out_json):
    """ncu csv with dram__bytes_{read,write}.sum + gpu__time_duration.sum + tensor pipe % for every conv launch of a
    step -> profiles/conv_traffic.json (mean DRAM bytes per launch, consumed by bench.py's roofline.traffic)."""
    import json
    with open(path) as handle:
        lines = [line for line in handle if not line.startswith('==')]
    rows = list(csv.DictReader(lines))
    per = collections.defaultdict(dict)
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for row in rows:
        value = float(row['Metric Value'].replace(',', ''))
        name = row['Metric Name']
        if name.startswith('dram__bytes'):
            value *= scale.get(row['Metric Unit'], 1.0)
        elif name == 'gpu__time_duration.sum':
            value = to_us(row['Metric Value'], row['Metric Unit'])
        per[row['ID']][name] = value
    n = len(per)
    read = sum(v.get('dram__bytes_read.sum', 0.0) for v in per.values())
    write = sum(v.get('dram__bytes_write.sum', 0.0) for v in per.values())
    time_us = sum(v.get('gpu__time_duration.sum', 0.0) for v in per.values())
    tensor = [v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') for v in per.values()]
    tensor_w = sum(t * v.get('gpu__time_duration.sum', 0.0) for t, v in zip(tensor, per.values()) if t is not None)
    result = {'source': path, 'launches': n, 'dram_bytes_per_launch': (read + write) / max(n, 1),
              'dram_read_bytes_total': read, 'dram_write_bytes_total': write, 'time_us_total': time_us,
              'tensor_pipe_active_pct_time_weighted': tensor_w / max(time_us, 1e-9),
              'dram_gbs_while_running': (read + write) / max(time_us, 1e-9) / 1e3}
    with open(out_json, 'w') as handle:
        json.dump(result, handle, indent=1)
    print(json.dumps(result, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](*sys.argv[2:])
