"""`TopImages` / `TopImagesDataset` (`src/milannotations/datasets.py:20-292`), uint8-backed."""
import pathlib
from typing import Iterable, NamedTuple, Optional, Sequence, Tuple

import numpy
import torch

_BYTE_TO_PT = torch.tensor(1.0 / 255.0, dtype=torch.float64).to(torch.float32)  # renormalize.py:118-139


class TopImages(NamedTuple):
    """`src/milannotations/datasets.py:20-26` (+ `as_pil_image_grid` etc. omitted: visualisation)."""
    layer: str
    unit: int
    images: torch.Tensor
    masks: torch.Tensor


class TopImagesDataset(torch.utils.data.Dataset):
    """Top-activating images for individual units."""

    def __init__(self, root, name: Optional[str] = None, layers: Optional[Iterable] = None, device=None,
                 transform_images=None, transform_masks=None, display_progress: bool = True):
        del display_progress
        root = pathlib.Path(root)
        if not root.is_dir():
            raise FileNotFoundError(f'root directory not found: {root}')
        if layers is None:
            layers = [f.name for f in root.iterdir() if f.is_dir()]
        if not layers:
            raise ValueError('no layers given and root has no subdirectories')
        if name is None:
            name = f'{root.parent.name}/{root.name}'
        self.root = root
        self.name = name
        self.layers = layers = tuple(sorted(str(layer) for layer in layers))
        self.device = device
        self.transform_images = transform_images
        self.transform_masks = transform_masks
        self.images_by_layer, self.masks_by_layer, self.units_by_layer = {}, {}, {}
        for layer in layers:
            images_file = root / str(layer) / 'images.npy'
            masks_file = root / str(layer) / 'masks.npy'
            for file in (images_file, masks_file):
                if not file.exists():
                    raise FileNotFoundError(f'{layer} is missing {file.name}')
            images = numpy.load(images_file, mmap_mode='r')
            masks = numpy.load(masks_file, mmap_mode='r')
            for kind, array in (('images', images), ('masks', masks)):
                if array.ndim != 5:
                    raise ValueError(f'expected 5D {kind}, got {array.ndim}D in layer {layer}')
            if images.shape[:2] != masks.shape[:2]:
                raise ValueError(f'layer {layer} masks/images have different # unit/images: '
                                 f'{images.shape[:2]} vs. {masks.shape[:2]}')
            if images.shape[3:] != masks.shape[3:]:
                raise ValueError(f'layer {layer} masks/images have different height/width '
                                 f'{images.shape[3:]} vs. {masks.shape[3:]}')
            units_file = root / str(layer) / 'units.npy'
            if units_file.exists():
                units = torch.from_numpy(numpy.load(units_file))
                if units.dim() != 1:
                    raise ValueError(f'expected 1D units, got {units.dim()}D')
            else:
                units = torch.arange(len(images))
            self.images_by_layer[layer] = images
            self.masks_by_layer[layer] = masks
            self.units_by_layer[layer] = units
        self._index = [(layer, i) for layer in layers for i in range(len(self.images_by_layer[layer]))]

    def _convert(self, layer: str, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        images = torch.from_numpy(numpy.array(self.images_by_layer[layer][i])).float().mul(_BYTE_TO_PT)
        masks = torch.from_numpy(numpy.array(self.masks_by_layer[layer][i])).float()
        if self.device is not None:
            images, masks = images.to(self.device), masks.to(self.device)
        if self.transform_images is not None:
            images = self.transform_images(images)
        if self.transform_masks is not None:
            masks = self.transform_masks(masks)
        return images, masks

    def __getitem__(self, index: int) -> TopImages:
        layer, i = self._index[index]
        images, masks = self._convert(layer, i)
        return TopImages(layer=str(layer), unit=int(self.units_by_layer[layer][i].item()), images=images, masks=masks)

    def __len__(self) -> int:
        return len(self._index)

    def batch_u8(self, lo: int, hi: int, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """uint8 (hi-lo, k, 3, H, W) images and (hi-lo, k, 1, H, W) masks for samples [lo, hi) (engine fast path).

        Copies straight from the memory-mapped .npy files into `out` (reusable pinned tensors, see
        `alloc_batch_u8`) — one memcpy per contiguous run of units, no intermediate stack.
        """
        if self.transform_images is not None or self.transform_masks is not None:
            raise RuntimeError('batch_u8 bypasses transforms; use __getitem__')
        n = hi - lo
        if out is None:
            out = self.alloc_batch_u8(n)
        images, masks = out[0][:n], out[1][:n]
        images_np, masks_np = images.numpy(), masks.numpy()
        pos = lo
        while pos < hi:  # contiguous runs inside one layer
            layer, start = self._index[pos]
            run = 1
            while pos + run < hi and self._index[pos + run] == (layer, start + run):
                run += 1
            numpy.copyto(images_np[pos - lo:pos - lo + run], self.images_by_layer[layer][start:start + run])
            numpy.copyto(masks_np[pos - lo:pos - lo + run], self.masks_by_layer[layer][start:start + run],
                         casting='unsafe')
            pos += run
        return images, masks

    def alloc_batch_u8(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reusable (pinned when CUDA is present) staging tensors for `batch_u8`."""
        layer, _ = self._index[0]
        im_shape = self.images_by_layer[layer].shape[1:]
        mk_shape = self.masks_by_layer[layer].shape[1:]
        pin = torch.cuda.is_available()
        return (torch.empty((n, *im_shape), dtype=torch.uint8, pin_memory=pin),
                torch.empty((n, *mk_shape), dtype=torch.uint8, pin_memory=pin))

    def lookup(self, layer, unit: int) -> TopImages:
        layer = str(layer)
        if layer not in self.images_by_layer:
            raise KeyError(f'layer "{layer}" does not exist')
        if unit >= len(self.images_by_layer[layer]):
            raise KeyError(f'layer "{layer}" has no unit {unit}')
        images, masks = self._convert(layer, unit)
        return TopImages(layer=layer, unit=unit, images=images, masks=masks)

    def unit(self, index: int):
        layer, i = self._index[index]
        return str(layer), int(self.units_by_layer[layer][i].item())

    def units(self, indices: Sequence[int]):
        return tuple(self.unit(index) for index in indices)

    @property
    def k(self) -> int:
        assert len(self) > 0, 'empty dataset?'
        layer, _ = self._index[0]
        return self.images_by_layer[layer].shape[1]
