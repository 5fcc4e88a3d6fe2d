"""Facade of `src/milan/encoders.py`: the same constructor / `forward` / `map` surface, computed by the CUDA
engine. Only `PyramidConvEncoder('resnet101')` — the encoder of every shipped MILAN checkpoint
(`scripts/train_milan.py:29-32,88`) — is implemented natively."""
from typing import Any, Mapping, Optional, Tuple

import torch

from neuron_descriptions_b200 import synthetic

KIND_SPATIAL = 'spatial'
KIND_PYRAMID = 'pyramid'


class Encoder:
    """`src/milan/encoders.py:23-148` (abstract)."""

    feature_shape: Tuple[int, ...]

    def forward(self, images: torch.Tensor, masks: Optional[torch.Tensor] = None, **kwargs: Any) -> torch.Tensor:
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def map(self, dataset, mask: bool = True, image_index: int = 2, mask_index: int = 3, batch_size: int = 16,
            device=None, display_progress_as: Optional[str] = None, **_: Any):
        """`Encoder.map`, `src/milan/encoders.py:61-148`: featurise a whole dataset -> TensorDataset."""
        from torch.utils import data
        if device is not None:
            self.to(device)
        features = []
        for lo in range(0, len(dataset), batch_size):
            samples = [dataset[i] for i in range(lo, min(lo + batch_size, len(dataset)))]
            images = torch.stack([torch.as_tensor(s[image_index]) for s in samples])
            masks = torch.stack([torch.as_tensor(s[mask_index]) for s in samples]) if mask else None
            shape = images.shape
            flat_images = images.view(-1, *shape[-3:])
            flat_masks = masks.view(-1, *masks.shape[-3:]) if masks is not None else None
            out = self(flat_images, masks=flat_masks)
            features.append(out.view(*shape[:-3], *self.feature_shape).cpu())
        return data.TensorDataset(torch.cat(features))


class PyramidConvEncoder(Encoder):
    """`src/milan/encoders.py:243-351`, config 'resnet101' only. Bound to an engine by the owning `Decoder`."""

    def __init__(self, config: str = 'resnet50', **kwargs: Any):
        configs = PyramidConvEncoder.configs()
        if config not in configs:
            raise ValueError(f'encoder not supported: {config}')
        if config != 'resnet101':
            raise NotImplementedError(f'milan_b200 implements the resnet101 pyramid encoder natively; got {config!r}')
        self.config = config
        self.kwargs = dict(kwargs)
        self.kwargs.setdefault('pretrained', True)  # kept for checkpoint round-trips; nothing is downloaded
        _, self.layers, feature_size = configs[config]
        self.feature_shape = (feature_size,)
        self._engine = None

    def bind(self, engine) -> 'PyramidConvEncoder':
        self._engine = engine
        return self

    def to(self, device):
        return self

    def eval(self):
        return self

    def forward(self, images: torch.Tensor, masks: Optional[torch.Tensor] = None, normalize: bool = True,
                **_: Any) -> torch.Tensor:
        if self._engine is None:
            raise RuntimeError('encoder is not bound to a CUDA engine: call Decoder.to("cuda") first '
                               '(milan_b200 has no CPU path)')
        if not normalize:
            raise NotImplementedError('normalize=False is not supported by the fused stem')
        return self._engine.encode(images, masks)

    def properties(self) -> Mapping[str, Any]:
        return {'config': self.config, **self.kwargs}

    @staticmethod
    def configs():
        """Layer tables of `src/milan/encoders.py:326-351`."""
        layers = ('conv1', 'layer1', 'layer2', 'layer3', 'layer4')
        return {
            'alexnet': (None, ('features.0', 'features.3', 'features.6', 'features.8', 'features.10'), 1152),
            'resnet18': (None, layers, 1024),
            'resnet50': (None, layers, 3904),
            'resnet101': (None, layers, synthetic.FEATURE_SIZE),
        }


class SpatialConvEncoder(Encoder):
    """`src/milan/encoders.py:159-236`: not used by any shipped checkpoint; not implemented natively."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('SpatialConvEncoder is outside the describe-neurons hot path (SURVEY.md #2)')


def parse(key: str):
    return {Type.__name__: Type for Type in (SpatialConvEncoder, PyramidConvEncoder)}[key]


def key(encoder: Encoder) -> str:
    return type(encoder).__name__


def encoder(kind: str = KIND_PYRAMID, **kwargs: Any) -> Encoder:
    """`src/milan/encoders.py:371-391`."""
    if kind == KIND_SPATIAL:
        encoder_t = SpatialConvEncoder
    elif kind == KIND_PYRAMID:
        encoder_t = PyramidConvEncoder
    else:
        encoder_t = parse(kind)
    return encoder_t(**kwargs)
