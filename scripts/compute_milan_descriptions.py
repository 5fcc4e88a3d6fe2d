"""Compute MILAN descriptions for a given model/dataset pair on B200.

Drop-in for the reference's `scripts/compute_milan_descriptions.py` (same positional args, flags and CSV
output, `scripts/compute_milan_descriptions.py:14-72` there), running the describe path on the CUDA engine.
Under `torchrun --nproc-per-node N` the neurons are sharded over the N GPUs (contiguous ranges, one all-gather of
the captions' token ids at the end) and rank 0 writes the CSV. Two additions the reference does not have, both
optional: `--resume-dir DIR` keeps every rank's finished shard (token ids) in DIR so that a job that lost a rank can be
re-run without repeating the shards already done, and `MILAN_TIMING_JSON=path` makes rank 0 write the phase timings.

    python -m scripts.compute_milan_descriptions alexnet imagenet --data-dir DATA --milan base
"""
import argparse
import csv
import json
import os
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))

import torch  # noqa: E402
from torch import cuda  # noqa: E402

from neuron_descriptions_b200 import milan, milannotations, sharding  # noqa: E402
from neuron_descriptions_b200.milan import loaders as milan_loaders  # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser(description='compute milan descriptions')
    parser.add_argument('model', help='model architecture (e.g. alexnet)')
    parser.add_argument('dataset', help='dataset model trained on (e.g. imagenet)')
    parser.add_argument('--temperature', type=float, default=.2, help='pmi temperature (default: .2)')
    parser.add_argument('--beam-size', type=int, default=50, help='beam size to rerank (default: 50)')
    parser.add_argument('--data-dir', type=pathlib.Path, help='root dir for datasets (default: project data dir)')
    parser.add_argument('--results-dir', type=pathlib.Path,
                        help='root dir for final results (default: <project results dir> / descriptions)')
    parser.add_argument('--milan', default=milannotations.KEYS.BASE, help='milan model to use (default: base)')
    parser.add_argument('--milan-path', type=pathlib.Path, help='explicit checkpoint path (default: models dir)')
    parser.add_argument('--device', help='manually set device (default: guessed)')
    parser.add_argument('--resume-dir', type=pathlib.Path,
                        help='keep finished per-rank shards here and reuse them on a re-run (default: off)')
    args = parser.parse_args(argv)

    clock = time.perf_counter
    t_start = clock()
    world, rank, local_rank = sharding.init_distributed()
    # Same precedence as the reference (scripts/compute_milan_descriptions.py:39).
    device = (args.device or ('cuda' if world == 1 else f'cuda:{local_rank}')) if cuda.is_available() else 'cpu'
    if not str(device).startswith('cuda'):
        raise SystemExit('milan_b200 is CUDA-only (B200, sm_100a): no CPU fallback; got device ' + str(device))

    key = f'{args.model}/{args.dataset}'
    data_dir = args.data_dir or milannotations.loaders.data_dir()
    data_root = data_dir / key
    results_dir = args.results_dir
    if results_dir is None:
        results_root = os.environ.get('MILAN_RESULTS_DIR')
        results_dir = (pathlib.Path(results_root) if results_root else
                       pathlib.Path(__file__).resolve().parents[1] / 'results') / 'descriptions'
    results_dir.mkdir(exist_ok=True, parents=True)

    decoder = milan_loaders.pretrained(args.milan, path=args.milan_path)
    t_checkpoint = clock()
    decoder.to(device)
    t_engine = clock()
    dataset = milannotations.load(key, path=data_root)

    t_loaded = clock()
    sharding.barrier(device)  # every rank starts its shard together: the describe phase is the max over ranks
    t_describe = clock()
    predictions = sharding.predict_sharded(decoder, dataset, world=world, rank=rank, strategy='rerank',
                                           temperature=args.temperature, beam_size=args.beam_size, device=device,
                                           resume_dir=args.resume_dir)
    cuda.synchronize(device)
    t_described = clock()
    if rank == 0:
        rows = [('layer', 'unit', 'description')]
        for index, description in enumerate(predictions):
            layer, unit = dataset.unit(index)
            rows.append((str(layer), str(unit), description))
        results_csv_file = results_dir / f'{key.replace("/", "_")}.csv'
        with results_csv_file.open('w') as handle:
            csv.writer(handle).writerows(rows)
        print(f'wrote {len(rows) - 1} descriptions to {results_csv_file}')
        timing_path = os.environ.get('MILAN_TIMING_JSON')
        if timing_path:
            with open(timing_path, 'w') as handle:
                json.dump({'neurons': len(dataset), 'world': world, 'load_s': t_loaded - t_start,
                           'load_checkpoint_s': t_checkpoint - t_start, 'build_engine_s': t_engine - t_checkpoint,
                           'open_dataset_s': t_loaded - t_engine,
                           'describe_s': t_described - t_describe, 'csv_s': clock() - t_described,
                           'total_s': clock() - t_start,
                           'rank0_predict': getattr(decoder, 'last_predict_timing', None)}, handle)
    sharding.finalize_distributed()


if __name__ == '__main__':
    main()
