"""BASELINE.json config 3 — "resnet152/places365 4k neurons, k=15, beam=50 + PMI rerank, 8xB200 neuron-sharded" —
measured as STRONG scaling through the reference's own entry point.

The reference's resnet152/places365 exemplar set covers conv1 + layer1..4 = 64 + 256 + 512 + 1024 + 2048 = 3904
units (`src/exemplars/models.py:321-326`); `scripts/compute_milan_descriptions.py:52-72` describes all of them. No
dataset or checkpoint is available offline, so `prepare` writes a synthetic exemplar set of exactly that shape in the
on-disk layout `TopImagesDataset` reads (`<root>/data/resnet152/places365/<layer>/{images,masks}.npy`, uint8) plus a
random-init MILAN checkpoint in the reference payload format, and `run` drives the unmodified CLI
(`scripts/compute_milan_descriptions.py`) over it:

    python scripts/config3_scaling.py prepare --root /dev/shm/milan_cfg3
    python scripts/config3_scaling.py run --root /dev/shm/milan_cfg3 --gpus 8      # spawns torchrun when --gpus > 1
    python scripts/config3_scaling.py compare --root /dev/shm/milan_cfg3 1 8       # CSVs of two runs must be identical

`run` prints ONE JSON line: wall-clock of the whole CLI process group (checkpoint load, engine build, mmap feed,
describe, all-gather, detokenise, CSV), and the describe phase alone (rank 0's `predict_sharded`, which ends with the
all-gather every rank joins, so it is the max over ranks), as neurons/s and neurons/hour.
"""
import argparse
import concurrent.futures
import json
import os
import pathlib
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

LAYERS = (('conv1', 64), ('layer1', 256), ('layer2', 512), ('layer3', 1024), ('layer4', 2048))
K = 15


def prepare(root: pathlib.Path, scale: float):
    import numpy as np
    from neuron_descriptions_b200 import milan, synthetic
    from neuron_descriptions_b200.milan import lang

    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer))
    decoder.load_state_dict(sd)
    (root / 'models').mkdir(parents=True, exist_ok=True)
    decoder.save(root / 'models' / 'base.pth')

    total = 0
    start = time.perf_counter()
    for li, (name, units) in enumerate(LAYERS):
        units = max(1, int(round(units * scale)))
        folder = root / 'data' / 'resnet152' / 'places365' / name
        folder.mkdir(parents=True, exist_ok=True)
        images = np.lib.format.open_memmap(folder / 'images.npy', mode='w+', dtype=np.uint8, shape=(units, K, 3, 224, 224))
        masks = np.lib.format.open_memmap(folder / 'masks.npy', mode='w+', dtype=np.uint8, shape=(units, K, 1, 224, 224))
        def fill(lo, li=li, units=units, images=images, masks=masks):
            n = min(64, units - lo)
            # masks as `synthetic.synthetic_exemplars` makes them (thresholded upsampled blobs, a few all-zero); image
            # bytes straight from numpy's generator, one generator per chunk so that the chunks can be made in parallel
            _, mk = synthetic.synthetic_exemplars(n, K, seed=5000 + 100 * li + lo // 64, images=False)
            rng = np.random.default_rng(7000 + 100 * li + lo // 64)
            images[lo:lo + n] = np.frombuffer(rng.bytes(n * K * 3 * 224 * 224), dtype=np.uint8).reshape(n, K, 3, 224, 224)
            masks[lo:lo + n] = mk.numpy()

        with concurrent.futures.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            list(pool.map(fill, range(0, units, 64)))
        images.flush()
        masks.flush()
        del images, masks
        total += units
    print(f'wrote {total} synthetic neurons x {K} exemplars under {root} in {time.perf_counter() - start:.1f} s')


def run(root: pathlib.Path, gpus: int, tag: str):
    results = root / f'results_{tag or gpus}'
    timing = results / 'timing.json'
    results.mkdir(parents=True, exist_ok=True)
    if timing.exists():
        timing.unlink()
    env = dict(os.environ, MILAN_MODELS_DIR=str(root / 'models'), MILAN_TIMING_JSON=str(timing), PYTHONPATH=str(ROOT))
    cli = [str(ROOT / 'scripts' / 'compute_milan_descriptions.py'), 'resnet152', 'places365', '--data-dir',
           str(root / 'data'), '--results-dir', str(results)]
    if gpus > 1:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={gpus}',
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + gpus)] + cli
    else:
        cmd = [sys.executable] + cli
    start = time.perf_counter()
    out = subprocess.run(cmd, env=env, capture_output=True, text=True)
    wall = time.perf_counter() - start
    if out.returncode != 0:
        sys.stderr.write(out.stdout[-4000:] + out.stderr[-4000:])
        raise SystemExit(f'CLI failed with rc {out.returncode}')
    with open(timing) as handle:
        phases = json.load(handle)
    n = phases['neurons']
    line = {
        'metric': 'neurons described/sec (k=15, beam=50), strong scaling over a fixed 3904-unit exemplar set',
        'config': {'workload': 'resnet152/places365 4k neurons (conv1 + layer1-4 = 3904 units), k=15, beam=50 + PMI rerank, '
                               'neuron-sharded; synthetic exemplars of that shape on disk, random-init MILAN weights',
                   'entry_point': 'scripts/compute_milan_descriptions.py resnet152 places365'},
        'n_gpus': gpus, 'neurons': n, 'scaling': 'strong',
        'describe_s': phases['describe_s'], 'value': n / phases['describe_s'], 'unit': 'neurons/s',
        'neurons_per_hour': 3600.0 * n / phases['describe_s'],
        'cli_wall_s': wall, 'neurons_per_hour_cli_wall': 3600.0 * n / wall,
        'phases_s': phases,
    }
    print(json.dumps(line), flush=True)


def compare(root: pathlib.Path, a: str, b: str):
    rows = []
    for tag in (a, b):
        with open(root / f'results_{tag}' / 'resnet152_places365.csv') as handle:
            rows.append(handle.read().splitlines())
    same = sum(x == y for x, y in zip(*rows))
    ok = len(rows[0]) == len(rows[1]) and same == len(rows[0])
    print(json.dumps({'compare': [a, b], 'rows': [len(r) - 1 for r in rows], 'identical_rows': same - 1, 'ok': ok}))
    if not ok:
        raise SystemExit(1)


def main():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('command', choices=('prepare', 'run', 'compare'))
    parser.add_argument('tags', nargs='*', help='compare: the two result tags')
    parser.add_argument('--root', type=pathlib.Path, default=pathlib.Path('/dev/shm/milan_cfg3'))
    parser.add_argument('--scale', type=float, default=1.0, help='prepare: fraction of the 3904 units (tests use less)')
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--tag', default='')
    args = parser.parse_intermixed_args()
    if args.command == 'prepare':
        prepare(args.root, args.scale)
    elif args.command == 'run':
        run(args.root, args.gpus, args.tag)
    else:
        compare(args.root, *args.tags)


if __name__ == '__main__':
    main()
