"""Dataset keys + `load` (`src/milannotations/loaders.py:11-87,227-268`), local files only."""
import os
import pathlib
import types

from neuron_descriptions_b200.milannotations import datasets

ENV_DATA_DIR = 'MILAN_DATA_DIR'

KEYS = types.SimpleNamespace(
    ALEXNET='alexnet', BIGGAN='biggan', DINO_VITS8='dino_vits8', RESNET152='resnet152', IMAGENET='imagenet',
    PLACES365='places365', ALEXNET_IMAGENET='alexnet/imagenet', ALEXNET_PLACES365='alexnet/places365',
    RESNET152_IMAGENET='resnet152/imagenet', RESNET152_PLACES365='resnet152/places365',
    BIGGAN_IMAGENET='biggan/imagenet', BIGGAN_PLACES365='biggan/places365',
    DINO_VITS8_IMAGENET='dino_vits8/imagenet', GENERATORS='gen', CLASSIFIERS='cls', BASE='base')


def data_dir() -> pathlib.Path:
    read = os.environ.get(ENV_DATA_DIR)
    if read is not None:
        return pathlib.Path(read)
    return pathlib.Path(__file__).resolve().parents[2] / 'data'


def load(name: str, path=None, **kwargs) -> datasets.TopImagesDataset:
    """Load the exemplar set `name` ('<model>/<dataset>') from `path` (default `$MILAN_DATA_DIR/<name>`)."""
    if path is None:
        path = data_dir() / name
    path = pathlib.Path(path)
    if not path.is_dir():
        raise FileNotFoundError(f'dataset path not found: {path} (downloads are disabled: no network)')
    kwargs.setdefault('name', name)
    return datasets.TopImagesDataset(path, **kwargs)
