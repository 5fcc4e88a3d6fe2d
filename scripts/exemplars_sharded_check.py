"""2+ ranks under torchrun (NCCL): stage-1 exemplars with the image dataset sharded over the GPUs must equal the
reference golden (and therefore the single-GPU result), in the exact and in the histogram regime."""
import os
import sys
import tempfile

import numpy as np
import torch
from torch.utils import data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuron_descriptions_b200 import exemplars, sharding  # noqa: E402
from oracle.make_golden import EXEMPLAR_CASES, exemplar_toy_images, exemplar_toy_model  # noqa: E402

world, rank, local_rank = sharding.init_distributed()
device = f'cuda:{local_rank}'
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'exemplars.npz'))
model, images = exemplar_toy_model(), exemplar_toy_images()
root = tempfile.mkdtemp() if rank == 0 else None
box = [root]
torch.distributed.broadcast_object_list(box, src=0)
root = box[0]
for layer, output_size, k in EXEMPLAR_CASES:
    stats = exemplars.discriminative(model, data.TensorDataset(images), layer=layer, device=device, results_dir=root,
                                     k=k, quantile=0.99, image_size=16, output_size=output_size, batch_size=8)
    assert stats.exact_quantile
    assert np.array_equal(stats.ids.cpu().numpy(), g[f'{layer}_ids']), (rank, layer)
    torch.distributed.barrier()
    if rank == 0:
        d = os.path.join(root, layer)
        assert np.array_equal(np.load(os.path.join(d, 'images.npy')), g[f'{layer}_images'])
        diff = int((np.load(os.path.join(d, 'masks.npy')) != g[f'{layer}_masks']).sum())
        assert diff <= 2, diff
        print(f'{layer}: sharded over {world} ranks == reference golden (mask pixels differing: {diff})')
# histogram regime: sharded == this rank alone on the whole dataset
big = torch.rand(24, 3, 64, 64, generator=torch.Generator().manual_seed(1))
sharded = exemplars.discriminative(model, data.TensorDataset(big), layer='conv_2', device=device, results_dir=None,
                                   save_results=False, k=5, quantile=0.99, output_size=64, batch_size=8)
assert not sharded.exact_quantile
torch.distributed.barrier()
torch.distributed.destroy_process_group()
single = exemplars.discriminative(model, data.TensorDataset(big), layer='conv_2', device=device, results_dir=None,
                                  save_results=False, k=5, quantile=0.99, output_size=64, batch_size=8)
assert torch.equal(sharded.ids.cpu(), single.ids.cpu()) and torch.equal(sharded.activations.cpu(), single.activations.cpu())
assert torch.equal(sharded.levels.cpu(), single.levels.cpu()), (sharded.levels, single.levels)
print(f'rank {rank}: histogram regime sharded == single GPU')
