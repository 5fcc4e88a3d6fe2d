"""`milan.pretrained()` — mirror of `src/milan/loaders.py` + the path logic of `src/utils/hubs.py:142-170` and
`src/utils/env.py:67-78`. Downloading is out of scope (no network): a missing file raises FileNotFoundError like
`hubs.ModelConfig.load` does when no URL is usable (`hubs.py:113-114`)."""
import os
import pathlib
from typing import Any

from neuron_descriptions_b200.milan import decoders

ENV_MODELS_DIR = 'MILAN_MODELS_DIR'
DEFAULT_MODELS_DIR = 'models'

# DATASET_GROUPINGS names not starting with NOT_ (src/milannotations/loaders.py:91-174), + '+clip' variants.
_GROUPS = ('base', 'cls', 'gen', 'imagenet', 'places365', 'alexnet', 'resnet152', 'biggan')
KEYS = tuple(_GROUPS) + tuple(f'{group}+clip' for group in _GROUPS)


def models_dir() -> pathlib.Path:
    read = os.environ.get(ENV_MODELS_DIR)
    if read is not None:
        return pathlib.Path(read)
    return pathlib.Path(__file__).resolve().parents[2] / DEFAULT_MODELS_DIR


def pretrained(config: str = 'base', path=None, **kwargs: Any) -> decoders.Decoder:
    """Return a pretrained MILAN model (`src/milan/loaders.py:28-32`)."""
    if config not in KEYS:
        raise KeyError(f'no such model in hub: {config}')
    if config.endswith('+clip'):
        raise NotImplementedError('CLIP reranking (DecoderWithCLIP) is out of scope (SURVEY.md #4)')
    if path is None:
        path = models_dir() / f'{config}.pth'
    path = pathlib.Path(path)
    if not path.exists():
        raise FileNotFoundError(f'model path not found: {path} (downloads are disabled: no network)')
    return decoders.Decoder.load(path, **kwargs)
