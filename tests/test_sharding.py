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
    resume = os.environ.get('MILAN_RESUME')
    if resume:  # per-rank shard persistence: a re-run only recomputes the shard whose file is missing
        assert sharding.predict_sharded(FakeDecoder(), dataset, world=world, rank=rank, resume_dir=resume) == expected
        lo, hi = sharding.shard_range(n, rank, world)
        saved = sharding._shard_file(resume, lo, hi, n, 4)
        assert saved.exists()
        class Exploding(FakeDecoder):
            def predict(self, dataset, batch_size=16, **kwargs):
                raise AssertionError('a finished shard was described again')
        if rank == 1:
            saved.unlink()  # "rank 1 died before finishing"
        decoder = FakeDecoder() if rank == 1 else Exploding()
        assert sharding.predict_sharded(decoder, dataset, world=world, rank=rank, resume_dir=resume) == expected
        assert saved.exists()
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
    env = dict(os.environ, MILAN_ROOT=ROOT, MILAN_N=str(n), CUDA_VISIBLE_DEVICES='',
               MILAN_RESUME=str(tmp_path / 'shards'))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('ok') == 2


def test_exemplar_select_topk_ties_and_empty_slots():
    from neuron_descriptions_b200.exemplars import sharding as es
    values = torch.tensor([[1.0, 3.0, 3.0, float('-inf'), 2.0, 3.0]])
    ids = torch.tensor([[7, 9, 4, -1, 1, 12]])
    v, i = es.select_topk(values, ids, 4)
    assert i.tolist() == [[4, 9, 12, 1]] and v.tolist() == [[3.0, 3.0, 3.0, 2.0]]
    v, i = es.select_topk(values[:, :1], torch.tensor([[-1]]), 2)
    assert i.tolist() == [[-1]] and v.tolist() == [[float('-inf')]]


EXEMPLAR_WORKER = textwrap.dedent('''
    import os, sys, torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ['MILAN_ROOT'])
    from neuron_descriptions_b200 import sharding as neuron_sharding
    from neuron_descriptions_b200.exemplars import sharding as es

    world, rank, _ = neuron_sharding.init_distributed()
    assert world == 2
    gen = torch.Generator().manual_seed(0)
    n, units, k, positions = 23, 5, 4, 6
    pooled = torch.randn(n, units, generator=gen)
    pooled[3, 1] = pooled[17, 1] = 9.0            # a tie across the two shards: the earlier image must win
    samples = torch.randn(n * positions, units, generator=gen)
    lo, hi = neuron_sharding.shard_range(n, rank, world)
    # what one rank's tally kernel would hold after its shard
    local_v, local_i = torch.topk(pooled[lo:hi].t(), k, dim=1)
    values, ids = es.merge_topk(local_v.contiguous(), (local_i + lo).contiguous())
    ref_v, ref_i = es.select_topk(pooled.t().contiguous(), torch.arange(n).repeat(units, 1), k)
    assert torch.equal(ids, ref_i) and torch.equal(values, ref_v), (rank, ids, ref_i)
    assert ids[1, 0].item() == 3 and ids[1, 1].item() == 17
    # exact-regime samples: concatenation in rank order, counts differ between the ranks
    cap = 256
    kept = torch.zeros(units, cap)
    mine = samples[lo * positions:hi * positions].t()
    kept[:, :mine.shape[1]] = mine
    merged, total = es.gather_samples(kept, mine.shape[1], cap)
    assert total == n * positions and torch.equal(merged[:, :total], samples.t())
    assert es.total_count(mine.shape[1], 'cpu') == n * positions
    # histograms add, byte results take the union
    hist = torch.zeros(units, 16, dtype=torch.int32)
    hist[:, rank] = rank + 1
    assert es.sum_histograms(hist)[:, :2].tolist() == [[1, 2]] * units
    masks = torch.zeros(4, dtype=torch.uint8)
    masks[rank * 2] = 1
    assert es.max_bytes(masks).tolist() == [1, 0, 1, 0]
    neuron_sharding.finalize_distributed()
    print('rank', rank, 'ok')
''')


def test_exemplar_statistics_merge_two_ranks_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(EXEMPLAR_WORKER)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MILAN_ROOT=ROOT, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('ok') == 2
