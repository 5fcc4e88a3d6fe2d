"""Dissect a vision model on B200: top-activating images + activation masks for every unit of the chosen layers.

Counterpart of the reference's `scripts/compute_exemplars.py` (same positional arguments and the flags that make
sense offline). The reference resolves `model/dataset` through a hub of downloadable weights and datasets
(`src/exemplars/models.py:160-403`, `src/exemplars/datasets.py:55-102`); there is no network here, so the model is
an architecture of `exemplars/models.py` (torchvision CNNs, DINO ViT-S/8) whose weights come from `--model-file` (a
`state_dict`), and the dataset is an image
folder at `--dataset-path` with the reference's ImageNet / Places365 transform (`datasets.py:60-75`). The results
(`<results-root>/<model>/<dataset>/<layer>/{images,masks}.npy, ids.csv, activations.csv`) are exactly what
`scripts/compute_milan_descriptions.py` reads; under `torchrun` the images shard over the GPUs.

    python -m scripts.compute_exemplars resnet18 imagenet --dataset-path /data/val --model-file resnet18.pth
"""
import argparse
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))

import torch  # noqa: E402
import torchvision  # noqa: E402
from torch import cuda  # noqa: E402

from neuron_descriptions_b200 import exemplars, sharding  # noqa: E402
from neuron_descriptions_b200.exemplars import models as zoo  # noqa: E402

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def main(argv=None):
    parser = argparse.ArgumentParser(description='compute unit exemplars')
    parser.add_argument('model', help='model architecture', choices=sorted(zoo.ZOO))
    parser.add_argument('dataset', help='dataset of unseen examples for model (names the results directory)')
    group = parser.add_mutually_exclusive_group()
    group.add_argument('--layer-names', nargs='+', help='layer names to compute exemplars for')
    group.add_argument('--layer-indices', type=int, nargs='+', help='indices into the model\'s default layers')
    parser.add_argument('--units', type=int, help='only compute exemplars for first n units (default: all)')
    parser.add_argument('--results-root', type=pathlib.Path,
                        help='exemplars results root (default: <project results dir> / exemplars)')
    parser.add_argument('--model-file', type=pathlib.Path, help='path to model weights (a state_dict)')
    parser.add_argument('--dataset-path', type=pathlib.Path, required=True, help='path to an image folder')
    parser.add_argument('--k', type=int, default=15, help='top images per unit (default: 15)')
    parser.add_argument('--quantile', type=float, default=.99, help='mask activation quantile (default: .99)')
    parser.add_argument('--batch-size', type=int, help='images per batch (default: the model entry\'s, else 128)')
    parser.add_argument('--device', help='manually set device (default: guessed)')
    args = parser.parse_args(argv)

    world, rank, local_rank = sharding.init_distributed()
    device = (args.device or ('cuda' if world == 1 else f'cuda:{local_rank}')) if cuda.is_available() else 'cpu'
    if not str(device).startswith('cuda'):
        raise SystemExit('milan_b200 is CUDA-only (B200, sm_100a): no CPU fallback; got device ' + str(device))

    model, layers, config = zoo.load(args.model, args.model_file)
    if args.model_file is None and rank == 0:
        print('warning: no --model-file given, dissecting a randomly initialised network', file=sys.stderr)
    dataset = torchvision.datasets.ImageFolder(
        str(args.dataset_path),
        transform=torchvision.transforms.Compose([
            torchvision.transforms.Resize(256), torchvision.transforms.CenterCrop(224),
            torchvision.transforms.ToTensor(), torchvision.transforms.Normalize(IMAGENET_MEAN, IMAGENET_STD)]))

    if args.layer_names:
        layers = args.layer_names
    elif args.layer_indices:
        layers = [layers[index] for index in args.layer_indices]
    results_root = args.results_root
    if results_root is None:
        base = os.environ.get('MILAN_RESULTS_DIR')
        results_root = (pathlib.Path(base) if base else pathlib.Path(__file__).resolve().parents[1] / 'results') / 'exemplars'
    results_dir = results_root / args.model / args.dataset
    for layer in layers:
        kwargs = dict(config, k=args.k, quantile=args.quantile, image_size=224, output_size=224)
        if args.batch_size:
            kwargs['batch_size'] = args.batch_size
        exemplars.discriminative(model, dataset, layer=layer, units=range(args.units) if args.units else None,
                                 results_dir=results_dir, device=device, **kwargs)
        if rank == 0:
            print(f'{args.model}/{args.dataset}/{layer}: exemplars written to {results_dir / str(layer)}')
    sharding.finalize_distributed()


if __name__ == '__main__':
    main()
