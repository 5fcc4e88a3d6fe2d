"""GPU: stage-1 exemplar computation (`neuron_descriptions_b200/exemplars`) against the golden produced by the
UNMODIFIED reference `exemplars.compute.discriminative` (`oracle/make_golden.py::make_exemplars_golden`) and
against the CPU oracle (`oracle/exemplars_oracle.py`) for the histogram regime and the kernels alone."""
import os

import numpy as np
import pytest
import torch
from torch.utils import data

from oracle import exemplars_oracle as E
from oracle.make_golden import (EXEMPLAR_CASES, GENERATIVE_CASES, VIT_CASE, ToyViT, exemplar_toy_images,
                                exemplar_toy_model, generative_toy_model)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('layer,output_size,k', EXEMPLAR_CASES)
def test_discriminative_matches_reference_golden(golden_dir, tmp_path, layer, output_size, k):
    from neuron_descriptions_b200 import exemplars
    g = np.load(os.path.join(golden_dir, 'exemplars.npz'))
    model, images = exemplar_toy_model(), exemplar_toy_images()
    stats = exemplars.discriminative(model, data.TensorDataset(images), layer=layer, device='cuda:0',
                                     results_dir=tmp_path, k=k, quantile=0.99, image_size=16,
                                     output_size=output_size, batch_size=8)
    assert stats.exact_quantile
    d = tmp_path / layer
    assert sorted(p.name for p in d.iterdir()) == ['activations.csv', 'ids.csv', 'images.npy', 'masks.npy']
    np.testing.assert_array_equal(np.loadtxt(d / 'ids.csv', delimiter=',').astype(np.int64), g[f'{layer}_ids'])
    np.testing.assert_allclose(np.loadtxt(d / 'activations.csv', delimiter=','), g[f'{layer}_activations'],
                               rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(np.load(d / 'images.npy'), g[f'{layer}_images'])
    got, ref = np.load(d / 'masks.npy'), g[f'{layer}_masks']
    assert got.shape == ref.shape and got.dtype == ref.dtype
    # the model forward runs on the GPU here and on the CPU in the reference: a pixel within an ulp of its level may flip
    assert int((got != ref).sum()) <= 2, f'{int((got != ref).sum())} mask pixels differ'
    # the files are a valid exemplar set for the describe path
    from neuron_descriptions_b200 import milannotations
    exemplar_set = milannotations.TopImagesDataset(tmp_path, layers=[layer])
    assert len(exemplar_set) == got.shape[0] and exemplar_set.k == k


def _check_result_dir(d, g, prefix, k, mask_slack=2):
    np.testing.assert_array_equal(np.loadtxt(d / 'ids.csv', delimiter=',').astype(np.int64), g[f'{prefix}ids'])
    np.testing.assert_allclose(np.loadtxt(d / 'activations.csv', delimiter=','), g[f'{prefix}activations'],
                               rtol=2e-5, atol=2e-6)
    got_images, ref_images = np.load(d / 'images.npy'), g[f'{prefix}images']
    assert got_images.shape == ref_images.shape and got_images.dtype == ref_images.dtype
    got, ref = np.load(d / 'masks.npy'), g[f'{prefix}masks']
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert int((got != ref).sum()) <= mask_slack, f'{int((got != ref).sum())} mask pixels differ'
    return got_images, ref_images


@pytest.mark.parametrize('layer,output_size,k,units', GENERATIVE_CASES)
def test_generative_matches_reference_golden(golden_dir, tmp_path, layer, output_size, k, units):
    """`exemplars.generative` vs the UNMODIFIED reference's (`src/exemplars/compute.py:352-437`): representations in,
    the generator's own output images kept for the top-k."""
    from neuron_descriptions_b200 import exemplars
    g = np.load(os.path.join(golden_dir, 'exemplars_generative.npz'))
    model, zs = generative_toy_model(), exemplar_toy_images(seed=11)
    stats = exemplars.generative(model, data.TensorDataset(zs), layer, device='cuda:0', results_dir=tmp_path, k=k,
                                 quantile=0.99, image_size=16, output_size=output_size, batch_size=8, units=units)
    assert stats.exact_quantile
    d = tmp_path / layer
    expected = ['activations.csv', 'ids.csv', 'images.npy', 'masks.npy'] + (['units.npy'] if units else [])
    assert sorted(p.name for p in d.iterdir()) == expected
    got_images, ref_images = _check_result_dir(d, g, f'{layer}_', k)
    # generated on the GPU here, on the CPU in the reference: sigmoid(x) * 255 truncated to a byte may land one off
    diff = np.abs(got_images.astype(np.int16) - ref_images.astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    if units:
        np.testing.assert_array_equal(np.load(d / 'units.npy'), g[f'{layer}_units'])


def test_vit_featurizer_matches_reference_golden(golden_dir, tmp_path):
    """The DINO ViT-S/8 configuration (`src/exemplars/models.py:236-247`): MLP activations of a ViT block, made
    spatial by `transforms.spatialize_vit_mlp`, through the same tally / mask kernels."""
    from neuron_descriptions_b200 import exemplars
    g = np.load(os.path.join(golden_dir, 'exemplars_vit.npz'))
    model, images = ToyViT().eval(), exemplar_toy_images(seed=23)
    layer, output_size, k = VIT_CASE
    stats = exemplars.discriminative(model, data.TensorDataset(images), layer=layer, device='cuda:0',
                                     results_dir=tmp_path, k=k, quantile=0.99, image_size=16, output_size=output_size,
                                     batch_size=8, transform_hiddens=exemplars.transforms.spatialize_vit_mlp)
    assert stats.exact_quantile
    got_images, ref_images = _check_result_dir(tmp_path / layer, g, '', k, mask_slack=4)
    np.testing.assert_array_equal(got_images, ref_images)


def test_tally_kernels_vs_oracle():
    import importlib
    C = importlib.import_module('neuron_descriptions_b200.exemplars.compute')
    gen = torch.Generator().manual_seed(3)
    hiddens = torch.randn(37, 9, 11, 13, generator=gen)
    hiddens[5, 2] = hiddens[9, 2]  # an exact tie between two images: the earlier index must win
    tally = C._Tally(5, torch.device('cuda:0'))
    for lo in range(0, 37, 10):
        tally.add(hiddens[lo:lo + 10].cuda())
    stats = tally.result(0.97)
    pooled, samples = E.pooled_and_samples(hiddens)
    values, ids = E.topk(pooled.numpy(), 5)
    np.testing.assert_array_equal(stats.ids.cpu().numpy(), ids)
    np.testing.assert_array_equal(stats.activations.cpu().numpy(), values)
    assert stats.exact_quantile and 37 * 11 * 13 <= E.EXACT_CAPACITY
    np.testing.assert_array_equal(stats.levels.cpu().numpy(), E.quantile_levels(samples.numpy(), np.float32(0.97)))


def test_histogram_quantile_beyond_the_exact_regime():
    import importlib
    C = importlib.import_module('neuron_descriptions_b200.exemplars.compute')
    gen = torch.Generator().manual_seed(4)
    hiddens = torch.randn(64, 7, 20, 20, generator=gen) * 3 + 1  # 25600 samples per unit > 8192
    tally = C._Tally(3, torch.device('cuda:0'))
    for lo in range(0, 64, 16):
        tally.add(hiddens[lo:lo + 16].cuda())
    stats = tally.result(0.99)
    assert not stats.exact_quantile
    samples = hiddens.permute(0, 2, 3, 1).reshape(-1, 7).numpy()
    n = len(samples)
    for u in range(7):
        s = np.sort(samples[:, u])
        exact = np.interp(0.99, (np.arange(n) + 0.5) / n, s)
        assert abs(stats.levels[u].item() - exact) <= abs(exact) * 2 ** -7, (stats.levels[u].item(), exact)


def test_activation_masks_vs_oracle():
    import importlib
    C = importlib.import_module('neuron_descriptions_b200.exemplars.compute')
    gen = torch.Generator().manual_seed(5)
    maps = torch.randn(6, 7, 7, generator=gen)
    levels = torch.tensor([0.0, 0.5, -0.5, 1.0, 0.2, 5.0])
    for size in (7, 16, 224):
        got = C.activation_masks(maps.cuda(), levels, size).cpu().numpy()
        for i in range(len(maps) if size < 100 else 1):
            ref = (E.upsample_bilinear_zeros(maps[i].numpy(), size) > levels[i].item()).astype(np.uint8)
            assert int((got[i] != ref).sum()) <= 1
    assert got[5].sum() == 0


def test_stage1_feeds_stage2(tmp_path):
    """Exemplars computed on the GPU (224x224, k = 3) are read back as a `TopImagesDataset` and described by the
    MILAN decoder: the two stages share one on-disk contract (`src/exemplars/compute.py:217-227` ->
    `src/milannotations/datasets.py:158-197`)."""
    from neuron_descriptions_b200 import exemplars, milan, milannotations, synthetic
    from neuron_descriptions_b200.milan import lang
    model = exemplar_toy_model()
    images = torch.rand(12, 3, 224, 224, generator=torch.Generator().manual_seed(8))
    stats = exemplars.discriminative(model, data.TensorDataset(images), layer='conv_2', device='cuda:0',
                                     results_dir=tmp_path, k=3, quantile=0.99, output_size=224, batch_size=4)
    assert not stats.exact_quantile  # 12 x 226 x 226 samples per unit: histogram regime
    exemplar_set = milannotations.TopImagesDataset(tmp_path, layers=['conv_2'])
    assert len(exemplar_set) == 6 and exemplar_set.k == 3
    sample = exemplar_set[0]
    assert sample.images.shape == (3, 3, 224, 224) and sample.masks.shape == (3, 1, 224, 224)
    assert 0.0 < float(sample.masks.mean()) < 0.05  # ~1 % of the pixels lie above the 0.99 quantile
    vocab = synthetic.synthetic_vocab(5000)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), max_neurons=16)
    decoder.load_state_dict(synthetic.synthetic_state_dict(seed=0, sharpen=12.0))
    captions = decoder.predict(exemplar_set, strategy='rerank', beam_size=10, device='cuda:0',
                               display_progress_as=None)
    assert len(captions) == 6 and all(isinstance(c, str) and c for c in captions)
