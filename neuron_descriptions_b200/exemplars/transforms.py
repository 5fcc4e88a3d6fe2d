"""Input / hidden transforms of the exemplar computation (mirror of `src/exemplars/transforms.py`)."""
import math
from typing import Any, Optional, Sequence, Tuple

import torch


def map_location(data: Sequence[Any], device: Optional[Any]) -> Tuple[Any, ...]:
    """`src/exemplars/transforms.py:11-29`: move every tensor of a batch to `device`."""
    return tuple(item.to(device) if isinstance(item, torch.Tensor) and device is not None else item for item in data)


def first(*inputs: Any) -> Tuple[Any, ...]:
    """`:37-39`: the first element of the batch is the model input."""
    return (inputs[0],)


def identity(inputs):
    """`:45-47`."""
    return inputs


def identities(*inputs):
    """`:50-52`."""
    return inputs


def spatialize_vit_mlp(hiddens: torch.Tensor) -> torch.Tensor:
    """`:55-81`: ViT MLP activations (batch, 1 + n_patches, units) -> (batch, units, side, side) without the CLS
    token, so that the CNN tally / mask kernels apply to a DINO ViT-S/8 (`dino_vits8`) unchanged."""
    batch_size, n_tokens, n_units = hiddens.shape
    side = math.isqrt(n_tokens - 1)
    if side * side != n_tokens - 1:
        raise AssertionError(f'{n_tokens - 1} patches do not form a square')
    return hiddens[:, 1:].permute(0, 2, 1).reshape(batch_size, n_units, side, side)
