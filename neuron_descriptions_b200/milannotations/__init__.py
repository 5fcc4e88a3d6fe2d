"""Exemplar-set input contract of the describe path: mirror of `src/milannotations` (loader side only).

SURVEY.md section 8(f) rank 1 ("next"): reads the on-disk format written by `src/exemplars/compute.py:217-227`
(`<root>/<layer>/{images,masks,units}.npy`, uint8) and presents `TopImagesDataset` samples with the reference's
value contract (`src/milannotations/datasets.py:157-197`: images float in [0,1] via the byte->pt renormalizer,
masks float). Unlike the reference it keeps the uint8 arrays memory-mapped and converts lazily, so a 4k-neuron
set costs 12 GB of page cache instead of 48 GB of fp32 RAM, and exposes `batch_u8` so the engine can ship uint8.
Annotations / merges / downloading are out of scope (training data plumbing).
"""
from neuron_descriptions_b200.milannotations.datasets import TopImages, TopImagesDataset
from neuron_descriptions_b200.milannotations.loaders import KEYS, load
