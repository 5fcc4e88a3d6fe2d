"""CPU restatement of the reference's stage-1 exemplar computation for a discriminative CNN.

TEST INFRASTRUCTURE ONLY (parity oracle for `neuron_descriptions_b200/exemplars`). It follows
`src/exemplars/compute.py:27-246,263-349` and the NetDissect pieces that path executes:

  * `RunningTopK` (`src/deps/netdissect/runningstats.py:31-118`): exact top-k of the spatially max-pooled
    activations per unit, with dataset indices; ties are implementation-defined in the reference (`torch.topk`),
    here the earlier dataset index wins;
  * `RunningQuantile` (`runningstats.py:274-420`) read out with `quantiles()` (`:557-580`): the sketch keeps EVERY
    sample until its first level (2 * r = 8192 columns, `tally.py:199-200` r = 4096) overflows, and its estimator
    is `numpy.interp(q, (cumsum(w) - w / 2) / sum(w), sorted samples)` with the minimum / maximum added as
    zero-weight end points. This restatement is exact in that regime only; beyond it the reference is a RANDOMISED
    KLL sketch with no deterministic output to pin;
  * `ImageVisualizer.pytorch_mask` / `pytorch_image` (`src/deps/netdissect/imgviz.py:185-211`) with the default
    grid of `upsample.upsample_grid` (`upsample.py:127-157`): `grid_sample(bilinear, zeros padding,
    align_corners=True)` at source position (t + 0.5) / scale - 0.5, thresholded with `>`; images are
    renormalised to bytes (`renormalize.py:118-139`) and resized with nearest-neighbour interpolation.

Pinned by `tests/golden/exemplars.npz`, produced by the unmodified reference (`oracle/make_golden.py`).
"""
from typing import Dict, Optional, Sequence

import numpy as np
import torch

EXACT_CAPACITY = 8192  # columns of the sketch's first level: 2 * r with r = 4096


def pooled_and_samples(hiddens: torch.Tensor):
    """`compute_topk_and_quantile`, `src/exemplars/compute.py:326-335`: (B,C,H,W) -> pooled (B,C), samples (B*H*W,C)."""
    batch, channels = hiddens.shape[:2]
    samples = hiddens.permute(0, 2, 3, 1).reshape(-1, channels)
    pooled = hiddens.reshape(batch, channels, -1).max(dim=2)[0]
    return pooled, samples


def topk(pooled: np.ndarray, k: int):
    """pooled (N, U) over the whole dataset -> values (U, k) descending, dataset indices (U, k)."""
    n, units = pooled.shape
    values = np.empty((units, k), np.float32)
    ids = np.empty((units, k), np.int64)
    for u in range(units):
        order = sorted(range(n), key=lambda i: (-pooled[i, u], i))[:k]
        ids[u] = order
        values[u] = pooled[order, u]
    return values, ids


def quantile_levels(samples: np.ndarray, q: float) -> np.ndarray:
    """samples (n, U) -> (U,) levels with the reference's estimator (exact regime)."""
    n, units = samples.shape
    assert n <= EXACT_CAPACITY, 'beyond the exact regime the reference sketch is randomised'
    levels = np.empty(units, np.float32)
    for u in range(units):
        s = np.sort(samples[:, u].astype(np.float32))
        xs = np.concatenate([[0.0], (np.arange(n, dtype=np.float32) + 0.5) / np.float32(n), [1.0]]).astype(np.float32)
        # cumsum(w) - w/2 over weights [0, 1...1, 0], divided by sum(w) = n, all in float32 like the reference
        w = np.concatenate([[0.0], np.ones(n, np.float32), [0.0]]).astype(np.float32)
        cw = (np.cumsum(w, dtype=np.float32) - w / 2) / np.float32(n)
        del xs
        ys = np.concatenate([[s[0]], s, [s[-1]]])
        levels[u] = np.float32(np.interp(q, cw, ys))
    return levels


def upsample_bilinear_zeros(act: np.ndarray, size: int) -> np.ndarray:
    """One (H, W) map -> (size, size) through the default NetDissect grid (see module docstring)."""
    h, w = act.shape
    out = np.zeros((size, size), np.float32)
    sy, sx = np.float32(size) / h, np.float32(size) / w
    for y in range(size):
        fy = (np.float32(y) - (np.float32(0.5) * sy - np.float32(0.5))) / sy if h > 1 else np.float32(0)
        for x in range(size):
            fx = (np.float32(x) - (np.float32(0.5) * sx - np.float32(0.5))) / sx if w > 1 else np.float32(0)
            y0, x0 = int(np.floor(fy)), int(np.floor(fx))
            wy1, wx1 = fy - y0, fx - x0
            acc = np.float32(0)
            for yy, wy in ((y0, 1 - wy1), (y0 + 1, wy1)):
                for xx, wx in ((x0, 1 - wx1), (x0 + 1, wx1)):
                    if 0 <= yy < h and 0 <= xx < w:
                        acc += np.float32(act[yy, xx]) * np.float32(wy) * np.float32(wx)
            out[y, x] = acc
    return out


def byte_images(images: torch.Tensor, size: int, mean: Sequence[float] = (0., 0., 0.),
                std: Sequence[float] = (1., 1., 1.)) -> torch.Tensor:
    """`pytorch_image`: undo the dataset normalisation into [0, 255] bytes, nearest-neighbour resize to `size`."""
    mul = torch.tensor(np.array(std) / np.array([1 / 255.] * 3)).to(images.dtype).view(1, 3, 1, 1)
    add = torch.tensor((np.array(mean) - 0.0) / np.array([1 / 255.] * 3)).to(images.dtype).view(1, 3, 1, 1)
    data = images.mul(mul).add_(add).clamp(0, 255).byte()
    return torch.nn.functional.interpolate(data.float(), size=(size, size)).clamp(0, 255).byte()


@torch.no_grad()
def discriminative(features_fn, images: torch.Tensor, k: int, quantile: float, output_size: int,
                   batch_size: int = 128, mean=(0., 0., 0.), std=(1., 1., 1.), images_fn=None,
                   units: Optional[Sequence[int]] = None) -> Dict[str, np.ndarray]:
    """`exemplars.compute.discriminative` for an in-memory image tensor; `features_fn(batch) -> (B,C,H,W)`.

    `images_fn` (generative models, `src/exemplars/compute.py:352-437`): the kept images are `images_fn(batch)` — the
    model's outputs — instead of the dataset items. `units` (`:159-175`): statistics of those channels only."""
    pooled, samples, hiddens = [], [], []
    shown = images if images_fn is None else torch.cat(
        [images_fn(images[lo:lo + batch_size]) for lo in range(0, len(images), batch_size)])
    for lo in range(0, len(images), batch_size):
        h = features_fn(images[lo:lo + batch_size])
        if units is not None:
            h = h[:, sorted(units)]
        p, s = pooled_and_samples(h)
        pooled.append(p.numpy())
        samples.append(s.numpy())
        hiddens.append(h.numpy())
    pooled, samples, hiddens = np.concatenate(pooled), np.concatenate(samples), np.concatenate(hiddens)
    values, ids = topk(pooled, k)
    levels = quantile_levels(samples, quantile)
    units = values.shape[0]
    masks = np.zeros((units, k, 1, output_size, output_size), np.uint8)
    top_images = np.zeros((units, k, 3, output_size, output_size), np.uint8)
    bytes_all = byte_images(shown, output_size, mean, std).numpy()
    for u in range(units):
        for r in range(k):
            up = upsample_bilinear_zeros(hiddens[ids[u, r], u], output_size)
            masks[u, r, 0] = (up > levels[u]).astype(np.uint8)
            top_images[u, r] = bytes_all[ids[u, r]]
    return {'ids': ids, 'activations': values, 'levels': levels, 'images': top_images, 'masks': masks}
