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
    with_reranker = config.endswith('+clip')
    if with_reranker:
        # `<group>+clip` (src/milan/loaders.py:13-25): the base checkpoint of the group with the CLIP reranker on top.
        # CLIP itself cannot be built offline, so the similarity model has to be handed in (milan/rerankers.py).
        if 'reranker' not in kwargs and 'similarity' not in (kwargs.get('reranker_kwargs') or {}):
            raise NotImplementedError(f'"{config}" needs CLIP ViT-B/32 (the `clip` package + weights), not available '
                                      'offline: pass reranker=rerankers.SimilarityReranker(<similarity callable>)')
        config = config[:-len('+clip')]
    if path is None:
        path = models_dir() / f'{config}.pth'
    path = pathlib.Path(path)
    if not path.exists():
        raise FileNotFoundError(f'model path not found: {path} (downloads are disabled: no network)')
    if with_reranker:
        extra = {key: kwargs.pop(key) for key in ('reranker', 'reranker_kwargs') if key in kwargs}
        return decoders.DecoderWithCLIP.from_decoder(decoders.Decoder.load(path, **kwargs), **extra)
    return decoders.Decoder.load(path, **kwargs)
