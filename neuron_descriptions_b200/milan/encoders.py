"""Facade of `src/milan/encoders.py`: the same constructor / `forward` / `map` surface, computed by the CUDA
engine. `PyramidConvEncoder('resnet101')` is the encoder of every shipped MILAN checkpoint
(`scripts/train_milan.py:29-32,88`); the other ResNet configs of the reference (`resnet18`, `resnet50` pyramids,
`SpatialConvEncoder('resnet18')`; `src/milan/encoders.py:214-216,326-351`) run on the same kernels with a
different layer table, and the `alexnet` pyramid runs its first convolution as an im2col GEMM."""
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

    def map(self, dataset, mask: bool = True, image_index: int = -3, mask_index: int = -2, batch_size: int = 64,
            device=None, display_progress_as: Optional[str] = None, **_: Any):
        """`Encoder.map`, `src/milan/encoders.py:61-148`: featurise a whole dataset -> TensorDataset."""
        from torch.utils import data
        if device is not None:
            self.to(device)
        features = []
        for lo in range(0, len(dataset), batch_size):
            samples = [dataset[i] for i in range(lo, min(lo + batch_size, len(dataset)))]
            # defaults (-3, -2) address AnnotatedTopImages (..., images, masks, annotations) like the reference's;
            # a plain TopImages sample needs image_index=2, mask_index=3 there too (encoders.py:63-64, :127-139)
            if not isinstance(samples[0][image_index], torch.Tensor):
                raise ValueError(f'non-tensor images: {type(samples[0][image_index]).__name__}')
            images = torch.stack([s[image_index] for s in samples])
            masks = None
            if mask:
                if not isinstance(samples[0][mask_index], torch.Tensor):
                    raise ValueError(f'non-tensor masks: {type(samples[0][mask_index]).__name__}')
                masks = torch.stack([s[mask_index] for s in samples])
            shape = images.shape
            flat_images = images.view(-1, *shape[-3:])
            flat_masks = masks.view(-1, *masks.shape[-3:]) if masks is not None else None
            out = self(flat_images, masks=flat_masks)
            features.append(out.view(*shape[:-3], *self.feature_shape).cpu())
        return data.TensorDataset(torch.cat(features))


class _EngineEncoder(Encoder):
    """Common part of the two encoder kinds: config validation, engine binding, standalone use.

    Inside a `Decoder` the encoder shares the decoder's engine (`bind`). Used on its own (as the reference's
    encoder tests do) it builds a private encoder-only engine from `load_state_dict` weights on `.to('cuda')`.
    """

    KIND = KIND_PYRAMID
    NATIVE = ()

    def __init__(self, config: str, **kwargs: Any):
        configs = type(self).configs()
        if config not in configs:
            raise ValueError(f'encoder not supported: {config}')
        if config not in self.NATIVE:
            raise NotImplementedError(f'milan_b200 implements {self.NATIVE} natively for {type(self).__name__}; '
                                      f'got {config!r}')
        self.config = config
        self.kwargs = dict(kwargs)
        self.kwargs.setdefault('pretrained', True)  # kept for checkpoint round-trips; nothing is downloaded
        self._engine = None
        self._own_engine = None
        self._state_dict = None
        self.max_images = 64

    def bind(self, engine) -> 'Encoder':
        self._engine = engine
        return self

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = False):
        """Weights with the reference's key names (`mean`, `std`, `encoder.model.*`: `nn.Module.state_dict()` of
        the reference encoder), for standalone use."""
        self._state_dict = {'encoder.' + key: value.detach().cpu() for key, value in state_dict.items()}
        if self._own_engine is not None:
            self._own_engine.close()
            self._own_engine = None
        return self

    def to(self, device):
        if device is None or self._state_dict is None:
            return self
        device = torch.device(device)
        if device.type != 'cuda':
            return self
        if self._own_engine is None or self._own_engine.device != torch.device('cuda', device.index or 0):
            from neuron_descriptions_b200.engine import Engine
            if self._own_engine is not None:
                self._own_engine.close()
            # encoder-only engine: the decoder dimensions are placeholders (no decoder weights are loaded)
            self._own_engine = Engine(self._state_dict, vocab_size=68, device=device,
                                      feature_size=self.feature_shape[-1], encoder_arch=self.config,
                                      encoder_kind=self.KIND, max_neurons=1, max_beam=1, max_keys=1,
                                      max_length=1, max_images=self.max_images, decoder=False)
        return self

    def eval(self):
        return self

    def forward(self, images: torch.Tensor, masks: Optional[torch.Tensor] = None, normalize: bool = True,
                **_: Any) -> torch.Tensor:
        engine = self._engine or self._own_engine
        if engine is None:
            raise RuntimeError('encoder is not bound to a CUDA engine: call Decoder.to("cuda") (or, standalone, '
                               'load_state_dict(...).to("cuda")) first; milan_b200 has no CPU path')
        if not normalize:
            raise NotImplementedError('normalize=False is not supported by the fused stem')
        return engine.encode(images, masks)

    def properties(self) -> Mapping[str, Any]:
        return {'config': self.config, **self.kwargs}


class PyramidConvEncoder(_EngineEncoder):
    """`src/milan/encoders.py:243-351`: masked spatial pooling of conv1 + layer1..4 -> one vector per image."""

    KIND = KIND_PYRAMID
    NATIVE = ('alexnet', 'resnet18', 'resnet50', 'resnet101')

    def __init__(self, config: str = 'resnet50', **kwargs: Any):
        super().__init__(config, **kwargs)
        _, self.layers, feature_size = self.configs()[config]
        self.feature_shape = (feature_size,)

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


class SpatialConvEncoder(_EngineEncoder):
    """`src/milan/encoders.py:159-236`: images * masks -> resnet18 layer4 -> (n, 49, 512)."""

    KIND = KIND_SPATIAL
    NATIVE = ('resnet18',)

    def __init__(self, config: str = 'resnet18', **kwargs: Any):
        super().__init__(config, **kwargs)
        _, layers, n_features, feature_size = self.configs()[config]
        self.layer, = layers
        self.feature_shape = (n_features, feature_size)

    def map(self, *args: Any, **kwargs: Any):
        """`SpatialConvEncoder.map` (`src/milan/encoders.py:218-226`): defaults for single-image datasets."""
        kwargs.setdefault('mask', False)
        kwargs.setdefault('image_index', 0)
        return super().map(*args, **kwargs)

    @staticmethod
    def configs():
        """`src/milan/encoders.py:232-236`."""
        return {'resnet18': (None, ('layer4',), 49, 512)}


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
