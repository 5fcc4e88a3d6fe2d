"""CPU: the exemplar-set loader mirrors the reference's input contract (tests/milannotations/datasets_test.py there):
on-disk layout, validation errors, images float in [0,1] via the byte->pt renormalizer, masks float, units.npy."""
import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import milannotations


def _write_layer(root, layer, n_units=3, k=5, size=16, units=None, seed=0):
    rng = np.random.RandomState(seed)
    layer_dir = root / layer
    layer_dir.mkdir(parents=True)
    images = rng.randint(0, 256, size=(n_units, k, 3, size, size), dtype=np.uint8)
    masks = rng.randint(0, 2, size=(n_units, k, 1, size, size), dtype=np.uint8)
    np.save(layer_dir / 'images.npy', images)
    np.save(layer_dir / 'masks.npy', masks)
    if units is not None:
        np.save(layer_dir / 'units.npy', np.asarray(units))
    return images, masks


def test_loads_layers_in_sorted_order_with_reference_value_contract(tmp_path):
    root = tmp_path / 'alexnet' / 'imagenet'
    im_b, mk_b = _write_layer(root, 'layer-b', seed=1)
    im_a, mk_a = _write_layer(root, 'layer-a', seed=2, units=[7, 8, 9])
    dataset = milannotations.TopImagesDataset(root)
    assert dataset.name == 'alexnet/imagenet'
    assert dataset.layers == ('layer-a', 'layer-b')
    assert len(dataset) == 6 and dataset.k == 5
    sample = dataset[0]
    assert isinstance(sample, milannotations.TopImages)
    assert (sample.layer, sample.unit) == ('layer-a', 7)
    assert sample.images.dtype == torch.float32 and sample.masks.dtype == torch.float32
    assert sample.images.shape == (5, 3, 16, 16) and sample.masks.shape == (5, 1, 16, 16)
    scale = torch.tensor(1.0 / 255.0, dtype=torch.float64).to(torch.float32)
    assert torch.equal(sample.images, torch.from_numpy(im_a[0]).float().mul(scale))  # renormalize.py:118-139
    assert sample.images.min() >= 0 and sample.images.max() <= 1
    assert torch.equal(sample.masks, torch.from_numpy(mk_a[0]).float())
    assert dataset.unit(4) == ('layer-b', 1)
    assert dataset.units([0, 5]) == (('layer-a', 7), ('layer-b', 2))
    looked = dataset.lookup('layer-b', 2)
    assert torch.equal(looked.masks, torch.from_numpy(mk_b[2]).float())
    with pytest.raises(KeyError):
        dataset.lookup('nope', 0)
    with pytest.raises(KeyError):
        dataset.lookup('layer-a', 99)
    images_u8, masks_u8 = dataset.batch_u8(2, 5)
    assert images_u8.dtype == torch.uint8 and images_u8.shape == (3, 5, 3, 16, 16)
    assert torch.equal(images_u8[0], torch.from_numpy(im_a[2])) and torch.equal(masks_u8[1], torch.from_numpy(mk_b[0]))


def test_validation_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        milannotations.TopImagesDataset(tmp_path / 'missing')
    empty = tmp_path / 'empty'
    empty.mkdir()
    with pytest.raises(ValueError, match='no layers'):
        milannotations.TopImagesDataset(empty)
    root = tmp_path / 'bad'
    (root / 'l0').mkdir(parents=True)
    np.save(root / 'l0' / 'images.npy', np.zeros((2, 3, 3, 8, 8), dtype=np.uint8))
    with pytest.raises(FileNotFoundError, match='masks.npy'):
        milannotations.TopImagesDataset(root)
    np.save(root / 'l0' / 'masks.npy', np.zeros((2, 3, 8, 8), dtype=np.uint8))
    with pytest.raises(ValueError, match='expected 5D masks'):
        milannotations.TopImagesDataset(root)
    np.save(root / 'l0' / 'masks.npy', np.zeros((2, 4, 1, 8, 8), dtype=np.uint8))
    with pytest.raises(ValueError, match='different # unit/images'):
        milannotations.TopImagesDataset(root)
    np.save(root / 'l0' / 'masks.npy', np.zeros((2, 3, 1, 8, 9), dtype=np.uint8))
    with pytest.raises(ValueError, match='different height/width'):
        milannotations.TopImagesDataset(root)
    with pytest.raises(FileNotFoundError):
        milannotations.load('alexnet/imagenet', path=tmp_path / 'nowhere')


def test_load_by_key(tmp_path):
    root = tmp_path / 'resnet152' / 'places365'
    _write_layer(root, 'conv1')
    dataset = milannotations.load(milannotations.KEYS.RESNET152_PLACES365, path=root)
    assert dataset.name == 'resnet152/places365' and len(dataset) == 3
