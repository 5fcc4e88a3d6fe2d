"""Top-activating images + activation masks for every unit of a layer: the step that PRODUCES the
`images.npy` / `masks.npy` exemplar sets the describe path consumes (`src/exemplars/compute.py:27-246,263-349`).

Same call surface as the reference for discriminative models (`discriminative(model, dataset, layer=..., k=15,
quantile=.99, output_size=224, ...)`, `compute(compute_topk_and_quantile, compute_activations, dataset, ...)`) and
the same files in `results_dir` (`images.npy`, `masks.npy`, `ids.csv`, `activations.csv`, `units.npy`). The
network being described runs as ordinary PyTorch on the GPU (a library call, like in the reference); the
statistics over its activations — running top-k, quantile, upsample + threshold masks — are this repo's CUDA
kernels behind the C ABI (`csrc/exemplars.cu`, `include/milan_b200.h`). No CPU fallback.

Differences from the reference, by design:
  * the quantile sketch. NetDissect's `RunningQuantile` keeps every sample while a unit has seen <= 8192 of them
    and is a RANDOMISED KLL sketch beyond; here the first regime is reproduced exactly and the second is a
    deterministic 65536-bin histogram of the float bit pattern (level error <= one bf16 ulp, 0.8 %);
  * visualisations (`viz_dir`, lightbox html, per-image PNGs) and the tally / mask cache files are not written.
"""
import ctypes
import pathlib
from typing import Any, Callable, NamedTuple, Optional, Sequence

import numpy
import torch
from torch.utils import data

from neuron_descriptions_b200 import _lib
from neuron_descriptions_b200 import sharding as neuron_sharding
from neuron_descriptions_b200.exemplars import sharding

EXACT_CAPACITY = 8192  # columns of the reference sketch's first level (2 * r, r = 4096: tally.py:199-200)


class ActivationStats(NamedTuple):
    """What the tally found: top-k pooled activations / dataset indices per unit, and the quantile levels."""
    activations: torch.Tensor  # (units, k) float32, descending
    ids: torch.Tensor          # (units, k) int64 dataset indices
    levels: torch.Tensor       # (units,) float32 activation level at `quantile`
    exact_quantile: bool       # True: the reference's exact regime; False: histogram read-out


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr())


def _check(rc: int, what: str):
    if rc != 0:
        raise _lib.MilanError(f'{what} failed with CUDA error {rc}')


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


from neuron_descriptions_b200.exemplars.transforms import first, identities, identity  # noqa: E402,F401


class _Tally:
    """Running top-k + quantile state on the device (RunningTopK + RunningQuantile of the reference)."""

    def __init__(self, k: int, device, first_index: int = 0):
        self.k, self.device = k, device
        self.lib = _lib.load()
        self.count = first_index  # dataset index of the next image (a rank's shard starts at `first_index`)
        self.samples_seen = 0   # activations per unit seen
        self.top_vals = self.top_ids = self.samples = self.hist = None
        self.kept = []          # batches kept while still in the exact regime

    def add(self, hiddens: torch.Tensor):
        hiddens = hiddens.to(self.device, torch.float32).contiguous()
        B, U = hiddens.shape[:2]
        P = hiddens[0, 0].numel()
        acts = hiddens.view(B, U, P)
        if self.top_vals is None:
            self._init_state(U)
        for lo in range(0, B, 1024):
            part = acts[lo:lo + 1024]
            pooled = torch.empty(len(part), U, device=self.device)
            _check(self.lib.milan_tally_topk(_ptr(part), len(part), U, P, self.count + lo, self.k, _ptr(pooled),
                                             _ptr(self.top_vals), _ptr(self.top_ids), _stream(self.device)),
                   'milan_tally_topk')
        if self.hist is None and self.samples_seen + B * P <= EXACT_CAPACITY:
            _check(self.lib.milan_tally_samples(_ptr(acts), B, U, P, _ptr(self.samples), EXACT_CAPACITY,
                                                self.samples_seen, _stream(self.device)), 'milan_tally_samples')
        else:
            if self.hist is None:  # leaving the exact regime: fold what was kept into the histogram
                self._to_histogram()
            _check(self.lib.milan_tally_hist(_ptr(acts), B, U, P, _ptr(self.hist), _stream(self.device)),
                   'milan_tally_hist')
        self.count += B
        self.samples_seen += B * P

    def _init_state(self, units: int):
        self.top_vals = torch.full((units, self.k), float('-inf'), device=self.device)
        self.top_ids = torch.full((units, self.k), -1, dtype=torch.long, device=self.device)
        self.samples = torch.empty(units, EXACT_CAPACITY, device=self.device)

    def _to_histogram(self):
        """Leave the exact regime: fold the kept samples into a fresh histogram."""
        U = self.samples.shape[0]
        self.hist = torch.zeros(U, 65536, dtype=torch.int32, device=self.device)
        if self.samples_seen:
            kept = self.samples[:, :self.samples_seen].t().contiguous().view(self.samples_seen, U, 1)
            _check(self.lib.milan_tally_hist(_ptr(kept), self.samples_seen, U, 1, _ptr(self.hist),
                                             _stream(self.device)), 'milan_tally_hist')
        self.samples = None

    def result(self, quantile: float) -> ActivationStats:
        """Read-out; with torch.distributed initialised the per-rank statistics are merged first (every rank gets
        the same answer; see `exemplars/sharding.py`)."""
        world, _ = sharding.world_and_rank()
        # a rank whose image shard was empty never ran add(): it learns the unit count from its peers and joins the
        # collectives below with empty statistics instead of leaving them blocked
        U = sharding.max_over_ranks(0 if self.top_vals is None else self.top_vals.shape[0], self.device)
        if U == 0:
            raise ValueError('no activations were tallied on any rank')
        if self.top_vals is None:
            self._init_state(U)
        top_vals, top_ids = sharding.merge_topk(self.top_vals, self.top_ids)
        total = sharding.total_count(self.samples_seen, self.device)
        # every rank must take the same branch: exact iff ALL samples of all ranks fit the sketch's first level
        if world > 1:
            flag = torch.tensor([1 if self.hist is None else 0], device=self.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            exact = bool(flag.item()) and total <= EXACT_CAPACITY
        else:
            exact = self.hist is None
        levels = torch.empty(U, device=self.device)
        if exact:
            samples, n = sharding.gather_samples(self.samples, self.samples_seen, EXACT_CAPACITY)
            _check(self.lib.milan_quantile_exact(_ptr(samples), U, EXACT_CAPACITY, n, float(quantile), _ptr(levels),
                                                 _stream(self.device)), 'milan_quantile_exact')
        else:
            if self.hist is None:
                self._to_histogram()
            hist = sharding.sum_histograms(self.hist.clone() if world > 1 else self.hist)
            _check(self.lib.milan_quantile_hist(_ptr(hist), U, total, float(quantile), _ptr(levels),
                                                _stream(self.device)), 'milan_quantile_hist')
        return ActivationStats(top_vals, top_ids, levels, exact)


def activation_masks(maps: torch.Tensor, levels: torch.Tensor, size: int) -> torch.Tensor:
    """(n, H, W) activation maps + (n,) levels -> (n, size, size) uint8 masks (`pytorch_mask`, imgviz.py:185-198)."""
    maps = maps.to(torch.float32).contiguous()
    levels = levels.to(maps.device, torch.float32).contiguous()
    n, H, W = maps.shape
    masks = torch.empty(n, size, size, dtype=torch.uint8, device=maps.device)
    _check(_lib.load().milan_activation_masks(_ptr(maps), _ptr(levels), n, H, W, size, _ptr(masks),
                                              _stream(maps.device)), 'milan_activation_masks')
    return masks


def _byte_images(images: torch.Tensor, size: int, mul, add) -> torch.Tensor:
    """`ImageVisualizer.pytorch_image` (imgviz.py:200-210): the renormaliser's map into bytes
    (`Renormalizer.__call__`, renormalize.py:129-139), nearest-neighbour resize to the output size."""
    mul = torch.tensor(mul).to(images.device, images.dtype).view(1, 3, 1, 1)
    add = torch.tensor(add).to(images.device, images.dtype).view(1, 3, 1, 1)
    data_ = images.mul(mul).add_(add).clamp(0, 255).byte()
    return torch.nn.functional.interpolate(data_.float(), size=(size, size)).clamp(0, 255).byte()


def _find_normalizer(source):
    """`renormalize.find_normalizer` (renormalize.py:95-115): the Normalize transform of a dataset, if any."""
    from torchvision import transforms
    if source is None:
        return None
    if isinstance(source, transforms.Normalize):
        return source
    t = getattr(source, 'transform', None)
    if t is not None:
        return _find_normalizer(t)
    for t in reversed(getattr(source, 'transforms', None) or ()):
        found = _find_normalizer(t)
        if found is not None:
            return found
    return None


def _byte_map(normalizer):
    """(mul, add) of `renormalize.renormalizer(source, target='byte')` (renormalize.py:53-80,118-127): data * mul +
    add is in [0, 255]. `normalizer` is a torchvision `Normalize` / anything with the SOURCE `(mean, std)`, a NetDissect
    `Renormalizer` that targets bytes (its own `mul` / `add` are used as they are), or None = 'pt' data in [0, 1]."""
    if normalizer is not None and hasattr(normalizer, 'mul') and hasattr(normalizer, 'add'):
        if not getattr(normalizer, 'tobyte', True):
            raise ValueError('renormalizer must target bytes (renormalize.renormalizer(..., target="byte"))')
        return (numpy.asarray(torch.as_tensor(normalizer.mul).flatten().double()),
                numpy.asarray(torch.as_tensor(normalizer.add).flatten().double()))
    if normalizer is None:
        mean, std = (0., 0., 0.), (1., 1., 1.)
    elif hasattr(normalizer, 'mean') and hasattr(normalizer, 'std'):
        mean, std = normalizer.mean, normalizer.std
    else:
        mean, std = normalizer
    byte_scale = numpy.array([1.0 / 255] * 3)
    return numpy.array(std, dtype=numpy.float64) / byte_scale, numpy.array(mean, dtype=numpy.float64) / byte_scale


def compute(compute_topk_and_quantile: Callable[..., torch.Tensor],
            compute_activations: Callable[..., torch.Tensor],
            dataset: data.Dataset,
            units: Optional[Sequence[int]] = None,
            k: int = 15,
            quantile: float = 0.99,
            output_size: int = 224,
            batch_size: int = 128,
            image_size: Optional[int] = None,
            renormalizer=None,
            num_workers: int = 0,
            results_dir=None,
            save_results: bool = True,
            device='cuda',
            display_progress: bool = True,
            **_: Any) -> ActivationStats:
    """`exemplars.compute` (`src/exemplars/compute.py:27-246`). Both callables take a dataset batch and return the
    layer's activations (B, C, H, W) on the device (the reference's first callable returns the pooled / flattened
    pair instead; pooling and flattening happen in the tally kernels here). `compute_activations` may return an
    `(activations, images)` pair instead (`:168-175`): the images to keep are then the model's outputs (generative
    models), not the dataset items."""
    del image_size, num_workers, display_progress
    if units is not None and not units:
        raise ValueError('when setting `units`, must provide >= 1 unit')
    if k < 1:
        raise ValueError(f'must have k >= 1, got k={k}')
    if quantile <= 0 or quantile >= 1:
        raise ValueError(f'must have quantile in range (0, 1), got quantile={quantile}')
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('milan_b200 exemplar statistics are CUDA-only (no CPU fallback)')
    if results_dir is not None:
        results_dir = pathlib.Path(results_dir)
        if save_results:
            results_dir.mkdir(exist_ok=True, parents=True)
    unit_index = None
    if units is not None:
        units = sorted(units)
        unit_index = torch.tensor(units, device=device)
        if save_results and results_dir is not None:
            numpy.save(f'{results_dir}/units.npy', numpy.array(units))

    def select(hiddens):
        hiddens = hiddens.to(device)
        return hiddens if unit_index is None else hiddens[:, unit_index]

    # ---- pass 1: tally (tally.tally_topk_and_quantile, tally.py:199-222); under torch.distributed every rank
    # tallies a contiguous range of the images and the statistics are merged in `_Tally.result`
    world, rank = sharding.world_and_rank()
    lo, hi = neuron_sharding.shard_range(len(dataset), rank, world)
    tally = _Tally(k, device, first_index=lo)
    loader = data.DataLoader(data.Subset(dataset, range(lo, hi)), batch_size=batch_size, shuffle=False)
    for batch in loader:
        batch = batch if isinstance(batch, (list, tuple)) else [batch]
        tally.add(select(compute_topk_and_quantile(*batch)))
    stats = tally.result(quantile)
    if not save_results and results_dir is None:
        return stats

    # ---- pass 2: masks + images of the top-k (ImageVisualizer.individual_masked_images_for_topk via
    # tally.gather_topk, tally.py:92-124): re-run the model on the needed images only
    ids = stats.ids.cpu()
    n_units = ids.shape[0]
    needed = sorted(image for image in set(ids.view(-1).tolist()) if image >= 0)
    needed = needed[slice(*neuron_sharding.shard_range(len(needed), rank, world))]  # this rank's share of pass 2
    owned = set(needed)
    normalizer = renormalizer if renormalizer is not None else _find_normalizer(dataset)
    to_byte = _byte_map(normalizer)
    masks = torch.zeros(n_units, k, 1, output_size, output_size, dtype=torch.uint8)
    images = torch.zeros(n_units, k, 3, output_size, output_size, dtype=torch.uint8)
    subset = data.Subset(dataset, needed)
    offset = 0
    wanted = {}  # image -> [(unit, rank)]
    for unit in range(n_units):
        for slot, image in enumerate(ids[unit].tolist()):
            if image in owned:
                wanted.setdefault(image, []).append((unit, slot))
    for batch in data.DataLoader(subset, batch_size=batch_size, shuffle=False):
        batch = batch if isinstance(batch, (list, tuple)) else [batch]
        outputs = compute_activations(*batch)
        shown = batch[0]
        if not torch.is_tensor(outputs):  # generative: (activations, generated images)
            outputs, shown = outputs
        hiddens = select(outputs).float()
        bytes_ = _byte_images(shown.to(device).float(), output_size, *to_byte).cpu()
        pairs, maps, levels = [], [], []
        for local in range(len(hiddens)):
            image = needed[offset + local]
            for unit, slot in wanted[image]:
                pairs.append((unit, slot, local))
                maps.append(hiddens[local, unit])
                levels.append(stats.levels[unit])
        got = activation_masks(torch.stack(maps), torch.stack(levels), output_size).cpu()
        for (unit, slot, local), mask in zip(pairs, got):
            masks[unit, slot, 0] = mask
            images[unit, slot] = bytes_[local]
        offset += len(hiddens)
    if world > 1:  # every slot was filled by exactly one rank
        masks = sharding.max_bytes(masks.to(device)).cpu()
        images = sharding.max_bytes(images.to(device)).cpu()
    if save_results and results_dir is not None and rank == 0:
        numpy.save(f'{results_dir}/images.npy', images.numpy())
        numpy.save(f'{results_dir}/masks.npy', masks.numpy())
        for metadata, name, fmt in ((stats.activations, 'activations', '%.5e'), (stats.ids, 'ids', '%i')):
            numpy.savetxt(str(results_dir / f'{name}.csv'), metadata.view(n_units, k).cpu().numpy(), delimiter=',',
                          fmt=fmt)
    return stats


def discriminative(model: torch.nn.Module,
                   dataset: data.Dataset,
                   layer: Optional[str] = None,
                   device='cuda',
                   results_dir=None,
                   viz_dir=None,
                   transform_inputs: Callable[..., Any] = first,
                   transform_hiddens: Callable[[torch.Tensor], torch.Tensor] = identity,
                   **kwargs: Any) -> ActivationStats:
    """`exemplars.discriminative` (`src/exemplars/compute.py:263-349`): exemplars of `layer` (default: the model's
    output) of an image classifier; results go to `<results_dir>/<layer or 'outputs'>`."""
    del viz_dir
    model.to(device).eval()
    if results_dir is not None:
        results_dir = pathlib.Path(results_dir) / (str(layer) if layer is not None else 'outputs')
    retained = {}
    handle = None
    if layer is not None:
        modules = dict(model.named_modules())
        if str(layer) not in modules:
            raise KeyError(f'layer "{layer}" not found in model')
        handle = modules[str(layer)].register_forward_hook(
            lambda _module, _inputs, output: retained.__setitem__('x', output))

    def activations(*args: Any) -> torch.Tensor:
        inputs = transform_inputs(*[a.to(device) if torch.is_tensor(a) else a for a in args])
        with torch.no_grad():
            outputs = model(**inputs) if isinstance(inputs, dict) else model(*inputs)
        hiddens = outputs if layer is None else retained['x']
        return transform_hiddens(hiddens)

    try:
        return compute(activations, activations, dataset, results_dir=results_dir, device=device, **kwargs)
    finally:
        if handle is not None:
            handle.remove()


def generative(model: torch.nn.Module,
               dataset: data.Dataset,
               layer: str,
               device='cuda',
               results_dir=None,
               viz_dir=None,
               transform_inputs: Callable[..., Any] = identities,
               transform_hiddens: Callable[[Any], torch.Tensor] = identity,
               transform_outputs: Callable[[Any], torch.Tensor] = identity,
               **kwargs: Any) -> ActivationStats:
    """`exemplars.generative` (`src/exemplars/compute.py:352-437`): exemplars of `layer` of a model for which a
    representation goes in and an image comes out (BigGAN: `dataset` holds the z / class pairs). The unit statistics
    are tallied over the layer's activations exactly as for a classifier; the images kept for the top-k are the
    model's OUTPUTS (through `transform_outputs`), renormalised to bytes (`renormalizer=`; default: [0, 1] data).
    Results go to `<results_dir>/<layer>`. The whole batch is handed to the model (`transform_inputs` default
    `identities`), as in the reference."""
    del viz_dir
    model.to(device).eval()
    if results_dir is not None:
        results_dir = pathlib.Path(results_dir) / str(layer)
    modules = dict(model.named_modules())
    if str(layer) not in modules:
        raise KeyError(f'layer "{layer}" not found in model')
    retained = {}
    handle = modules[str(layer)].register_forward_hook(
        lambda _module, _inputs, output: retained.__setitem__('x', output))

    def run(*args: Any):
        inputs = transform_inputs(*[a.to(device) if torch.is_tensor(a) else a for a in args])
        with torch.no_grad():
            images = model(**inputs) if isinstance(inputs, dict) else model(*inputs)
        return transform_hiddens(retained['x']), images

    def tallied(*args: Any) -> torch.Tensor:
        return run(*args)[0]

    def activations_and_images(*args: Any):
        hiddens, images = run(*args)
        return hiddens, transform_outputs(images)

    try:
        return compute(tallied, activations_and_images, dataset, results_dir=results_dir, device=device, **kwargs)
    finally:
        handle.remove()
