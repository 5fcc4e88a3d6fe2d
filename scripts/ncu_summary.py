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




def traffic(path, out_json, kernel='conv_gemm', limit=100):
    """ncu csv with dram__bytes_{read,write}.sum + gpu__time_duration.sum + tensor pipe % for every conv launch of a
    step -> profiles/conv_traffic.json (mean DRAM bytes per launch, consumed by bench.py's roofline.traffic)."""
    import json
    with open(path) as handle:
        lines = [line for line in handle if not line.startswith('==')]
    rows = [row for row in csv.DictReader(lines) if kernel in row['Kernel Name']]
    keep = list(collections.OrderedDict((row['ID'], None) for row in rows))[:int(limit)]  # the encoder's launches
    rows = [row for row in rows if row['ID'] in set(keep)]
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


def layers(path, out):
    """Same csv as `traffic`: one line per encoder conv launch of the resnet101 step, in execution order
    (stem, then conv1 / conv2 / conv3 of every bottleneck; the x.0 blocks run conv3 + downsample as one launch)."""
    with open(path) as handle:
        lines = [line for line in handle if not line.startswith('==')]
    per = collections.OrderedDict()
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for row in csv.DictReader(lines):
        if 'conv_gemm' not in row['Kernel Name']:  # conv_gemm_kernel<...> and conv_gemm_pair_kernel
            continue
        value = float(row['Metric Value'].replace(',', ''))
        name = row['Metric Name']
        if name.startswith('dram__bytes'):
            value *= scale.get(row['Metric Unit'], 1.0)
        elif name == 'gpu__time_duration.sum':
            value = to_us(row['Metric Value'], row['Metric Unit'])
        per.setdefault(row['ID'], {})[name] = value
    names = ['stem 7x7/2 3->64']
    for li, (blocks, planes) in enumerate(((3, 64), (4, 128), (23, 256), (3, 512)), start=1):
        for b in range(blocks):
            stride = '/2' if (b == 0 and li > 1) else ''
            names.append(f'layer{li}.{b}.conv1 1x1 ->{planes}')
            names.append(f'layer{li}.{b}.conv2 3x3{stride} ->{planes}')
            names.append(f'layer{li}.{b}.conv3' + ('+downsample' if b == 0 else '') + f' 1x1 ->{planes * 4}'
                         + ('' if b == 0 else ' +res'))
    rows = list(per.values())[:len(names)]
    tensor_key = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
    with open(out, 'w') as handle:
        handle.write('# ncu --metrics gpu__time_duration.sum,' + tensor_key + ',dram__bytes_read.sum,dram__bytes_write.sum '
                     '--clock-control none\n# one bench step (64 neurons = 960 images): the encoder conv launches in '
                     f'execution order ({path})\n')
        handle.write(f'{"conv":44s} {"us":>9s} {"tensor%":>8s} {"rd MB":>8s} {"wr MB":>8s} {"DRAM GB/s":>10s}\n')
        tot = collections.Counter()
        for name, v in zip(names, rows):
            t = v.get('gpu__time_duration.sum', 0.0)
            rd, wr = v.get('dram__bytes_read.sum', 0.0), v.get('dram__bytes_write.sum', 0.0)
            handle.write(f'{name:44s} {t:9.1f} {v.get(tensor_key, 0.0):8.1f} {rd / 1e6:8.0f} {wr / 1e6:8.0f} '
                         f'{(rd + wr) / max(t, 1e-9) / 1e3:10.0f}\n')
            tot['t'] += t; tot['rd'] += rd; tot['wr'] += wr; tot['tw'] += t * v.get(tensor_key, 0.0)
        handle.write(f'{"TOTAL":44s} {tot["t"]:9.1f} {tot["tw"] / max(tot["t"], 1e-9):8.1f} {tot["rd"] / 1e6:8.0f} '
                     f'{tot["wr"] / 1e6:8.0f} {(tot["rd"] + tot["wr"]) / max(tot["t"], 1e-9) / 1e3:10.0f}\n')
    print(open(out).read()[-400:])


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic, 'layers': layers}[sys.argv[1]](*sys.argv[2:])
