"""The networks the stage-1 CLI can dissect offline (counterpart of `src/exemplars/models.py:160-403`).

The reference resolves `<model>/<dataset>` through a hub of downloadable weights (torchvision zoo, `torch.hub.load(
'facebookresearch/dino:main', 'dino_vits8')`, a BigGAN zoo). None of that is reachable here, so every entry builds
the ARCHITECTURE with random initial weights and takes real weights from `--model-file` (a `state_dict`). The network
being described is ordinary PyTorch — a library call in the reference too; what this repo accelerates is the
statistics over its activations (`exemplars/compute.py`) and the describe path that consumes the result.
"""
from typing import Any, Callable, Dict, NamedTuple, Optional, Sequence

import torch
import torchvision
from torch import nn

from neuron_descriptions_b200.exemplars import transforms


class _VitAttention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        q, k, v = self.qkv(x).view(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        return self.proj(nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c))


class _VitMlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _VitBlock(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1, self.attn = nn.LayerNorm(dim, eps=1e-6), _VitAttention(dim, heads)
        self.norm2, self.mlp = nn.LayerNorm(dim, eps=1e-6), _VitMlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, patch: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """DINO's `vit_small(patch_size=8)` (ViT-S/8: 384 wide, 12 blocks, 6 heads, MLP 1536) with the parameter names of
    facebookresearch/dino `vision_transformer.py`, so that its `dino_deitsmall8_pretrain.pth` state_dict loads. The
    units the reference dissects are the MLP hidden layers `blocks.<i>.mlp.fc1` (`models.py:46`), shape
    (batch, 1 + 784, 1536), made spatial by `transforms.spatialize_vit_mlp`. Fixed 224 x 224 inputs (the reference's
    dataset transform); DINO's position-embedding interpolation for other sizes is not needed."""

    def __init__(self, image_size: int = 224, patch: int = 8, dim: int = 384, depth: int = 12, heads: int = 6,
                 mlp_ratio: float = 4.0):
        super().__init__()
        self.patch_embed = _PatchEmbed(patch, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + (image_size // patch) ** 2, dim))
        self.blocks = nn.ModuleList(_VitBlock(dim, heads, mlp_ratio) for _ in range(depth))
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)

    def forward(self, images):
        x = self.patch_embed(images)
        x = torch.cat([self.cls_token.expand(len(x), -1, -1), x], dim=1) + self.pos_embed
        for block in self.blocks:
            x = block(x)
        return self.norm(x)[:, 0]


class ModelEntry(NamedTuple):
    factory: Callable[[], nn.Module]
    layers: Sequence[str]
    kwargs: Dict[str, Any]  # forwarded to exemplars.discriminative (the reference's ModelExemplarsConfig.kwargs)


def _torchvision(name: str) -> Callable[[], nn.Module]:
    return lambda: getattr(torchvision.models, name)(weights=None)


_RESNET_LAYERS = ('conv1', 'layer1', 'layer2', 'layer3', 'layer4')
# default layers as in `src/exemplars/models.py:33-60` LAYERS
ZOO: Dict[str, ModelEntry] = {
    'alexnet': ModelEntry(_torchvision('alexnet'), ('features.0', 'features.3', 'features.6', 'features.8', 'features.10'), {}),
    'resnet18': ModelEntry(_torchvision('resnet18'), _RESNET_LAYERS, {}),
    'resnet34': ModelEntry(_torchvision('resnet34'), _RESNET_LAYERS, {}),
    'resnet50': ModelEntry(_torchvision('resnet50'), _RESNET_LAYERS, {}),
    'resnet101': ModelEntry(_torchvision('resnet101'), _RESNET_LAYERS, {}),
    'resnet152': ModelEntry(_torchvision('resnet152'), _RESNET_LAYERS, {}),
    # `models.py:236-247`: transform_hiddens=spatialize_vit_mlp, batch_size=32
    'dino_vits8': ModelEntry(VisionTransformer, tuple(f'blocks.{i}.mlp.fc1' for i in range(12)),
                             {'transform_hiddens': transforms.spatialize_vit_mlp, 'batch_size': 32}),
}


def load(name: str, model_file: Optional[str] = None):
    """(model, default layers, exemplar kwargs) for a zoo entry; weights from `model_file` if given."""
    if name not in ZOO:
        raise KeyError(f'unknown model "{name}" (known: {sorted(ZOO)}); generative models (BigGAN) have no offline '
                       'architecture here: call exemplars.generative(model, zs, layer) with your own generator')
    entry = ZOO[name]
    model = entry.factory()
    if model_file is not None:
        state = torch.load(model_file, map_location='cpu')
        model.load_state_dict(state.get('state_dict', state) if isinstance(state, dict) else state)
    return model.eval(), entry.layers, dict(entry.kwargs)
