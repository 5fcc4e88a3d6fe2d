"""CPU, world_size 2 over gloo: the N>1 host logic (contiguous neuron shards + one all-gather of token ids)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_dataset_in_order():
    for n in (0, 1, 7, 64, 1000, 3904):
        for world in (1, 2, 4, 8):
            pieces = [sharding.shard_range(n, r, world) for r in range(world)]
            flat = [i for lo, hi in pieces for i in range(lo, hi)]
            assert flat == list(range(n))
            per = (n + world - 1) // world
            assert all(hi - lo <= per for lo, hi in pieces)


WORKER = textwrap.dedent('''
    import os, sys, torch
    sys.path.insert(0, os.environ['MILAN_ROOT'])
    from neuron_descriptions_b200 import sharding
    from neuron_descriptions_b200.milan import lang

    class FakeDecoder:
        """Stands in for the CUDA decoder: "describes" neuron i as tokens [i % 7, i % 5, <stop>, ...]."""
        length = 4
        indexer = lang.Indexer(lang.Vocab(tuple(f'w{i}' for i in range(10))), start=True, stop=True, pad=True, unk=True)
        def predict(self, dataset, batch_size=16, **kwargs):
            ids = torch.tensor([[dataset[i] % 7, dataset[i] % 5, 11, 11] for i in range(len(dataset))],
                               dtype=torch.long).view(-1, 4)
            self.last_predict_tokens = ids
            return tuple(self.indexer.reconstruct(ids.tolist())) if len(ids) else ()

    world, rank, _ = sharding.init_distributed()
    assert world == 2
    n = int(os.environ['MILAN_N'])
    dataset = list(range(n))
    captions = sharding.predict_sharded(FakeDecoder(), dataset, world=world, rank=rank)
    expected = tuple(f'W{i % 7} w{i % 5}' for i in range(n))
    assert captions == expected, (rank, captions[:5], expected[:5])
    sharding.finalize_distributed()
    print('rank', rank, 'ok', len(captions))
''')


@pytest.mark.parametrize('n', [9, 32])
def test_predict_sharded_two_ranks_gloo(tmp_path, n):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MILAN_ROOT=ROOT, MILAN_N=str(n), CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('ok') == 2
