"""GPU: the `scripts/compute_milan_descriptions.py` entry point end to end on a tiny on-disk exemplar set and a
synthetic checkpoint in the reference payload format; captions are checked against the oracle's predict."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import milan, synthetic
from neuron_descriptions_b200.milan import lang
from oracle import milan_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compute_milan_descriptions_cli(tmp_path):
    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=1.0)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer))
    decoder.load_state_dict(sd)
    models = tmp_path / 'models'
    models.mkdir()
    decoder.save(models / 'base.pth')

    images_u8, masks_u8 = synthetic.synthetic_exemplars(5, 15, seed=21)
    data = tmp_path / 'data' / 'alexnet' / 'imagenet'
    for name, sl, units in (('conv1', slice(0, 2), [3, 9]), ('conv2', slice(2, 5), None)):
        (data / name).mkdir(parents=True)
        np.save(data / name / 'images.npy', images_u8[sl].numpy())
        np.save(data / name / 'masks.npy', masks_u8[sl].numpy())
        if units is not None:
            np.save(data / name / 'units.npy', np.asarray(units))
    results = tmp_path / 'results'
    env = dict(os.environ, MILAN_MODELS_DIR=str(models), PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'compute_milan_descriptions.py'), 'alexnet',
                          'imagenet', '--data-dir', str(tmp_path / 'data'), '--results-dir', str(results),
                          '--beam-size', '10'], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    with open(results / 'alexnet_imagenet.csv') as handle:
        rows = list(csv.reader(handle))
    assert rows[0] == ['layer', 'unit', 'description']
    assert [r[:2] for r in rows[1:]] == [['conv1', '3'], ['conv1', '9'], ['conv2', '0'], ['conv2', '1'], ['conv2', '2']]
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        ref = O.describe(images_f, masks_f, sd, vocab, strategy='rerank', beam_size=10, temperature=.2)
    assert [r[2] for r in rows[1:]] == list(ref)


def test_predict_streams_chunks_from_disk(tmp_path):
    """Decoder.predict over an on-disk TopImagesDataset with more neurons than one engine chunk: exercises the
    prefetching uint8 host feed (two pinned staging buffers) and per-chunk grouping."""
    from neuron_descriptions_b200 import milannotations
    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=1.0)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), max_neurons=2)
    decoder.load_state_dict(sd)
    images_u8, masks_u8 = synthetic.synthetic_exemplars(5, 15, seed=33)
    root = tmp_path / 'resnet152' / 'places365'
    for name, sl in (('layer1', slice(0, 3)), ('layer2', slice(3, 5))):
        (root / name).mkdir(parents=True)
        np.save(root / name / 'images.npy', images_u8[sl].numpy())
        np.save(root / name / 'masks.npy', masks_u8[sl].numpy())
    dataset = milannotations.load('resnet152/places365', path=root)
    captions = decoder.predict(dataset, strategy='rerank', beam_size=8, batch_size=2, device='cuda:0',
                               display_progress_as=None)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        ref = O.describe(images_f, masks_f, sd, vocab, batch_size=2, strategy='rerank', beam_size=8)
    assert list(captions) == list(ref)
    assert decoder.last_predict_tokens.shape == (5, 15)
    # the same through precomputed features (Encoder.map + predict(features=...), src/milan/encoders.py:61-148)
    feats = decoder.encoder.map(dataset, image_index=2, mask_index=3, batch_size=2, device='cuda:0',
                               display_progress_as=None)
    with pytest.raises(ValueError, match='non-tensor images'):  # reference defaults (-3, -2) hit `unit` here
        decoder.encoder.map(dataset, batch_size=2, display_progress_as=None)
    again = decoder.predict(dataset, features=feats, strategy='rerank', beam_size=8, batch_size=2,
                            display_progress_as=None)
    assert list(again) == list(ref)


def test_compute_exemplars_cli(tmp_path):
    """`scripts/compute_exemplars.py` (stage 1) on an image folder -> files `TopImagesDataset` reads."""
    from PIL import Image
    from scripts import compute_exemplars
    from neuron_descriptions_b200 import milannotations
    rng = np.random.default_rng(0)
    for cls in ('a', 'b'):
        (tmp_path / 'images' / cls).mkdir(parents=True)
        for i in range(5):
            Image.fromarray(rng.integers(0, 256, (260, 300, 3), dtype=np.uint8)).save(tmp_path / 'images' / cls / f'{i}.png')
    compute_exemplars.main(['resnet18', 'toyset', '--dataset-path', str(tmp_path / 'images'), '--results-root',
                            str(tmp_path / 'exemplars'), '--layer-names', 'layer4', '--units', '4', '--k', '3',
                            '--batch-size', '4', '--device', 'cuda:0'])
    root = tmp_path / 'exemplars' / 'resnet18' / 'toyset'
    assert sorted(p.name for p in (root / 'layer4').iterdir()) == ['activations.csv', 'ids.csv', 'images.npy',
                                                                    'masks.npy', 'units.npy']
    dataset = milannotations.TopImagesDataset(root, layers=['layer4'])
    assert len(dataset) == 4 and dataset.k == 3
    sample = dataset[1]
    assert sample.unit == 1 and sample.images.shape == (3, 3, 224, 224) and sample.masks.shape == (3, 1, 224, 224)
    ids = np.loadtxt(root / 'layer4' / 'ids.csv', delimiter=',')
    assert ids.shape == (4, 3) and ids.min() >= 0 and ids.max() < 10
    # the DINO ViT-S/8 entry (`src/exemplars/models.py:236-247`): MLP units of a block on the 28 x 28 patch grid
    compute_exemplars.main(['dino_vits8', 'toyset', '--dataset-path', str(tmp_path / 'images'), '--results-root',
                            str(tmp_path / 'exemplars'), '--layer-names', 'blocks.1.mlp.fc1', '--units', '5', '--k', '2',
                            '--device', 'cuda:0'])
    dataset = milannotations.TopImagesDataset(tmp_path / 'exemplars' / 'dino_vits8' / 'toyset', layers=['blocks.1.mlp.fc1'])
    assert len(dataset) == 5 and dataset.k == 2 and dataset[0].masks.shape == (2, 1, 224, 224)
    assert 0 < float(dataset[0].masks.float().mean()) < 0.2  # 0.99-quantile masks: a few percent of the pixels


def test_bench_line_contract():
    """`python bench.py` (a short run): ONE JSON line with every key the driver reads, the roofline / e2e / clocks
    sub-objects, a non-zero count of this library's kernel launches and the same `config` dict as the reference arm."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '3', '--warmup', '3',
                          '--neurons-per-step', '16', '--no-cpu-baseline', '--no-fast-mode', '--no-strong-scaling'],
                         env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in line, key
    assert line['unit'] == 'neurons/s' and line['higher_is_better'] is True and line['scaling'] == 'weak'
    assert line['value'] > 0 and line['e2e']['value'] > 0 and line['gpu_launches'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 16 * 15 * 4 * 224 * 224 and line['e2e']['d2h_bytes_per_step'] > 0
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert key in line['roofline'], key
    assert line['roofline']['bound'] == 'tensor' and 0 < line['roofline']['frac'] < 1
    assert abs(line['roofline']['frac'] - line['roofline']['achieved'] / line['roofline']['peak']) < 1e-9
    sys.path.insert(0, ROOT)
    import bench
    assert line['config'] == bench.shared_config() and 'model' not in line['config']


def test_smoke_without_cta_pair_convs():
    """The cta_group::2 conv kernel is the default; MILAN_PAIR=0 (read once per process) selects the single-CTA
    kernel for the same convs. smoke() checks features, rerank scores and token ids against the oracle."""
    env = dict(os.environ, MILAN_PAIR='0', PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, '__graft_entry__.py'), 'smoke'], env=env,
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'tokens identical: True' in out.stdout
