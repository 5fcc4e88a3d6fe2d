"""Seeded synthetic MILAN checkpoints and exemplar sets (no network: real weights/data are unavailable).

`synthetic_state_dict` produces a flat state dict with exactly the key names of a reference `Decoder`
checkpoint (`src/utils/serialize.py:188-219`; key list in SURVEY.md section 5): the torchvision ResNet-101 of
`PyramidConvEncoder('resnet101')` under `encoder.encoder.model.*`, the attention-LSTM decoder, and the 2-layer
LSTM language model under `lm.*`. The same dict loads into the reference modules (`load_state_dict`) and into
this engine, which is how parity is checked.

`synthetic_exemplars` produces uint8 images and binary masks with the layout of MILANNOTATIONS
`images.npy` / `masks.npy` (`src/exemplars/compute.py:217-227`).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch

RESNET101_BLOCKS = (3, 4, 23, 3)
RESNET101_PLANES = (64, 128, 256, 512)
# torchvision ResNet variants: name -> (bottleneck?, blocks per stage)
RESNET_ARCHS = {'resnet101': (True, (3, 4, 23, 3)), 'resnet50': (True, (3, 4, 6, 3)),
                'resnet18': (False, (2, 2, 2, 2)), 'resnet34': (False, (3, 4, 6, 3))}
FEATURE_SIZE = 64 + 256 + 512 + 1024 + 2048  # 3904, src/milan/encoders.py:346-350
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # src/deps/netdissect/renormalize.py:87
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_vocab(size: int = 5000) -> Tuple[str, ...]:
    """A fake vocabulary; the first entries exercise punctuation handling in `Indexer.reconstruct`."""
    head = ['.', ',', '-', ';', ':', 'the', 'of', 'and', 'a', 'in', 'dog', 'cat', 'sky', 'grass', 'top', 'edges',
            'animals', 'objects', 'red', 'blue', 'round', 'text', 'faces', 'water', 'buildings', 'wheels']
    head = head[:size]
    return tuple(head + [f'w{i}' for i in range(len(head), size)])


def _conv(gen, cout, cin, k):
    std = math.sqrt(2.0 / (k * k * cout))  # torchvision: kaiming_normal_(mode='fan_out', relu)
    return torch.randn(cout, cin, k, k, generator=gen) * std


def _bn(gen, sd, prefix, c, gamma_lo, gamma_hi):
    sd[prefix + '.weight'] = torch.rand(c, generator=gen) * (gamma_hi - gamma_lo) + gamma_lo
    sd[prefix + '.bias'] = torch.randn(c, generator=gen) * 0.1
    sd[prefix + '.running_mean'] = torch.randn(c, generator=gen) * 0.1
    sd[prefix + '.running_var'] = torch.rand(c, generator=gen) + 0.5
    sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def synthetic_resnet_state(arch: str = 'resnet101', seed: int = 0) -> Dict[str, torch.Tensor]:
    """torchvision-ResNet-shaped weights with non-trivial BN statistics (so BN folding is exercised)."""
    bottleneck, stage_blocks = RESNET_ARCHS[arch]
    expansion = 4 if bottleneck else 1
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd['conv1.weight'] = _conv(gen, 64, 3, 7)
    _bn(gen, sd, 'bn1', 64, 0.5, 1.0)
    inplanes = 64
    for li, (blocks, planes) in enumerate(zip(stage_blocks, RESNET101_PLANES), start=1):
        for bi in range(blocks):
            pre = f'layer{li}.{bi}'
            stride = 2 if (bi == 0 and li > 1) else 1
            if bottleneck:
                sd[pre + '.conv1.weight'] = _conv(gen, planes, inplanes, 1)
                _bn(gen, sd, pre + '.bn1', planes, 0.5, 1.0)
                sd[pre + '.conv2.weight'] = _conv(gen, planes, planes, 3)
                _bn(gen, sd, pre + '.bn2', planes, 0.5, 1.0)
                sd[pre + '.conv3.weight'] = _conv(gen, planes * 4, planes, 1)
                _bn(gen, sd, pre + '.bn3', planes * 4, 0.1, 0.3)
            else:
                sd[pre + '.conv1.weight'] = _conv(gen, planes, inplanes, 3)
                _bn(gen, sd, pre + '.bn1', planes, 0.5, 1.0)
                sd[pre + '.conv2.weight'] = _conv(gen, planes, planes, 3)
                _bn(gen, sd, pre + '.bn2', planes, 0.2, 0.5)
            if stride != 1 or inplanes != planes * expansion:
                sd[pre + '.downsample.0.weight'] = _conv(gen, planes * expansion, inplanes, 1)
                _bn(gen, sd, pre + '.downsample.1', planes * expansion, 0.3, 0.6)
            inplanes = planes * expansion
    sd['fc.weight'] = torch.randn(1000, 512 * expansion, generator=gen) * 0.01  # unused by the encoders
    sd['fc.bias'] = torch.zeros(1000)
    return sd


def synthetic_alexnet_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    """torchvision-alexnet-shaped `features` weights (convs with biases, no BN); the classifier is omitted
    (the reference loads with strict=False and the pyramid encoder never reads it)."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, cout, cin, k in (('features.0', 64, 3, 11), ('features.3', 192, 64, 5), ('features.6', 384, 192, 3),
                               ('features.8', 256, 384, 3), ('features.10', 256, 256, 3)):
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=gen) * math.sqrt(2.0 / (k * k * cin))
        sd[name + '.bias'] = torch.randn(cout, generator=gen) * 0.1
    return sd


def synthetic_resnet101_state(seed: int = 0) -> Dict[str, torch.Tensor]:
    return synthetic_resnet_state('resnet101', seed)


def synthetic_encoder_state_dict(arch: str = 'resnet101', seed: int = 0) -> Dict[str, torch.Tensor]:
    """`state_dict()` of a reference encoder module: `mean`, `std`, `encoder.model.*`."""
    sd = {'mean': torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), 'std': torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)}
    backbone = synthetic_alexnet_state(seed) if arch == 'alexnet' else synthetic_resnet_state(arch, seed)
    for key, value in backbone.items():
        sd['encoder.model.' + key] = value
    return sd


def _linear(gen, sd, prefix, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    sd[prefix + '.weight'] = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    sd[prefix + '.bias'] = (torch.rand(out_f, generator=gen) * 2 - 1) * bound


def _lstm(gen, sd, prefix, suffix, in_f, hidden):
    bound = 1.0 / math.sqrt(hidden)
    sd[f'{prefix}weight_ih{suffix}'] = (torch.rand(4 * hidden, in_f, generator=gen) * 2 - 1) * bound
    sd[f'{prefix}weight_hh{suffix}'] = (torch.rand(4 * hidden, hidden, generator=gen) * 2 - 1) * bound
    sd[f'{prefix}bias_ih{suffix}'] = (torch.rand(4 * hidden, generator=gen) * 2 - 1) * bound
    sd[f'{prefix}bias_hh{suffix}'] = (torch.rand(4 * hidden, generator=gen) * 2 - 1) * bound


def synthetic_state_dict(seed: int = 0,
                         vocab_size: int = 5000,
                         embedding_size: int = 128,
                         hidden_size: int = 512,
                         sharpen: float = 3.0,
                         stop_bias: float = 0.0,
                         with_lm: bool = True,
                         feature_size: int = FEATURE_SIZE,
                         with_encoder: bool = True,
                         encoder_arch: str = 'resnet101') -> Dict[str, torch.Tensor]:
    """Full reference-format `Decoder` state dict.

    `sharpen` scales the vocab projection so top-k order is not rounding noise; `stop_bias` is added to the
    `<stop>` logit so beams terminate (exercises forced-stop and early exit). V = vocab_size + 4 specials
    (`src/utils/lang.py:242-260`).
    """
    gen = torch.Generator().manual_seed(seed + 1000)
    V = vocab_size + 4
    stop_index = vocab_size + 1
    H, E, F = hidden_size, embedding_size, feature_size
    A = min(H, F)  # src/milan/decoders.py:50
    sd: Dict[str, torch.Tensor] = {}
    if with_encoder:
        sd['encoder.mean'] = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        sd['encoder.std'] = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        for key, value in synthetic_resnet_state(encoder_arch, seed).items():
            sd['encoder.encoder.model.' + key] = value
    _linear(gen, sd, 'init_h.0', H, F)
    _linear(gen, sd, 'init_c.0', H, F)
    sd['embedding.weight'] = torch.randn(V, E, generator=gen)
    _linear(gen, sd, 'attend.query_to_hidden', A, H)
    _linear(gen, sd, 'attend.key_to_hidden', A, F)
    _linear(gen, sd, 'attend.output.0', 1, A)
    sd['attend.output.0.weight'] *= 8.0  # make attention non-uniform over the k exemplars
    _linear(gen, sd, 'feature_gate.0', F, H)
    _lstm(gen, sd, 'lstm.', '', E + F, H)
    _linear(gen, sd, 'output.1', V, H)
    sd['output.1.weight'] *= sharpen
    sd['output.1.bias'][stop_index] += stop_bias
    if with_lm:
        sd['lm.embedding.weight'] = torch.randn(V, E, generator=gen)
        sd['lm.embedding.weight'][vocab_size + 2].zero_()  # padding_idx row, src/milan/lms.py:47-49
        _lstm(gen, sd, 'lm.lstm.', '_l0', E, H)
        _lstm(gen, sd, 'lm.lstm.', '_l1', H, H)
        _linear(gen, sd, 'lm.output.0', V, H)
        sd['lm.output.0.weight'] *= sharpen
    return sd


def synthetic_exemplars(n_neurons: int,
                        k: int = 15,
                        size: int = 224,
                        seed: int = 0,
                        zero_mask_fraction: float = 0.02,
                        images: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """uint8 images (n, k, 3, size, size) and binary uint8 masks (n, k, 1, size, size).

    Masks look like thresholded upsampled activation maps (a few connected blobs covering a few percent of the
    image, cf. the 0.99-quantile masks of `src/exemplars/compute.py:32,195`); a fraction is all-zero to
    exercise `src/milan/encoders.py:311-314`.
    """
    gen = torch.Generator().manual_seed(seed + 77)
    # (images=False: masks only, drawn from a generator of their own — callers that make the image bytes themselves)
    images = (torch.randint(0, 256, (n_neurons, k, 3, size, size), generator=gen, dtype=torch.uint8) if images
              else None)
    low = torch.rand(n_neurons * k, 1, 14, 14, generator=gen)
    up = torch.nn.functional.interpolate(low, size=(size, size), mode='bilinear', align_corners=False)
    masks = (up > 0.8).to(torch.uint8).view(n_neurons, k, 1, size, size)
    drop = torch.rand(n_neurons, k, generator=gen) < zero_mask_fraction
    masks[drop] = 0
    return images, masks


def resnet101_conv_list() -> List[Tuple[str, int, int, int, int]]:
    """(name, cin, cout, ksize, stride) for the 104 convolutions through layer4, in execution order."""
    convs = [('conv1', 3, 64, 7, 2)]
    inplanes = 64
    for li, (blocks, planes) in enumerate(zip(RESNET101_BLOCKS, RESNET101_PLANES), start=1):
        for bi in range(blocks):
            stride = 2 if (bi == 0 and li > 1) else 1
            pre = f'layer{li}.{bi}'
            convs.append((pre + '.conv1', inplanes, planes, 1, 1))
            convs.append((pre + '.conv2', planes, planes, 3, stride))
            convs.append((pre + '.conv3', planes, planes * 4, 1, 1))
            if bi == 0:
                convs.append((pre + '.downsample.0', inplanes, planes * 4, 1, stride))
            inplanes = planes * 4
    return convs
