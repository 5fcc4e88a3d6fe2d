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


class _LayerExemplars(NamedTuple):
    """One layer directory as written by stage 1 (`src/exemplars/compute.py:217-227`), memory-mapped."""
    images: numpy.ndarray  # (units, k, 3, H, W) uint8
    masks: numpy.ndarray   # (units, k, 1, H, W) uint8
    units: torch.Tensor    # (units,) unit numbers

    @classmethod
    def open(cls, directory: pathlib.Path, layer: str) -> '_LayerExemplars':
        arrays = {}
        for kind in ('images', 'masks'):
            file = directory / f'{kind}.npy'
            if not file.exists():
                raise FileNotFoundError(f'{layer} is missing {file.name}')
            arrays[kind] = numpy.load(file, mmap_mode='r')  # stays on disk: 12 MB per unit as fp32 otherwise
        for kind, array in arrays.items():
            if array.ndim != 5:
                raise ValueError(f'expected 5D {kind}, got {array.ndim}D in layer {layer}')
        (n_img, k_img, _, *hw_img), (n_msk, k_msk, _, *hw_msk) = arrays['images'].shape, arrays['masks'].shape
        if (n_img, k_img) != (n_msk, k_msk):
            raise ValueError(f'layer {layer} masks/images have different # unit/images: '
                             f'{(n_img, k_img)} vs. {(n_msk, k_msk)}')
        if hw_img != hw_msk:
            raise ValueError(f'layer {layer} masks/images have different height/width '
                             f'{tuple(hw_img)} vs. {tuple(hw_msk)}')
        units = torch.arange(n_img)  # units.npy is optional: stage 1 writes it only for a unit subset
        if (directory / 'units.npy').exists():
            units = torch.from_numpy(numpy.load(directory / 'units.npy'))
            if units.dim() != 1:
                raise ValueError(f'expected 1D units, got {units.dim()}D')
        return cls(arrays['images'], arrays['masks'], units)


class TopImagesDataset(torch.utils.data.Dataset):
    """Top-activating images for individual units (`src/milannotations/datasets.py:93-292`).

    Same constructor, sample type, errors and value distribution as the reference (images = byte / 255 in fp32,
    masks 0.0 / 1.0), but the exemplar files stay memory-mapped uint8 and samples are converted on access; the
    engine reads them as uint8 through `batch_u8`.
    """

    def __init__(self, root, name: Optional[str] = None, layers: Optional[Iterable] = None, device=None,
                 transform_images=None, transform_masks=None, display_progress: bool = True):
        del display_progress  # nothing is read eagerly, so there is nothing to show progress for
        self.root = pathlib.Path(root)
        if not self.root.is_dir():
            raise FileNotFoundError(f'root directory not found: {self.root}')
        found = [entry.name for entry in self.root.iterdir() if entry.is_dir()] if layers is None else list(layers)
        if not found:
            raise ValueError('no layers given and root has no subdirectories')
        self.layers = tuple(sorted(map(str, found)))
        self.name = name if name is not None else f'{self.root.parent.name}/{self.root.name}'
        self.device = device
        self.transform_images, self.transform_masks = transform_images, transform_masks
        opened = {layer: _LayerExemplars.open(self.root / layer, layer) for layer in self.layers}
        self.images_by_layer = {layer: files.images for layer, files in opened.items()}
        self.masks_by_layer = {layer: files.masks for layer, files in opened.items()}
        self.units_by_layer = {layer: files.units for layer, files in opened.items()}
        # dataset order = layers sorted by name, units in file order (what the CSV of stage 2 follows)
        self._index = [(layer, row) for layer in self.layers for row in range(len(opened[layer].images))]

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
        """Row `unit` of `layer` (`datasets.py:238-259`; like the reference, `unit` is the row, not a unit number)."""
        stored = self.images_by_layer.get(str(layer))
        if stored is None:
            raise KeyError(f'layer "{layer}" does not exist')
        if not unit < len(stored):
            raise KeyError(f'layer "{layer}" has no unit {unit}')
        images, masks = self._convert(str(layer), unit)
        return TopImages(str(layer), unit, images, masks)

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
