"""Throughput of the stage-1 tally kernels on activation tensors of realistic shape (GPU box). They are single-pass
scans: the number to compare with is the HBM copy bandwidth (MEASURED_PEAKS.json)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device('cuda', 0)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
for name, (B, U, H) in {'resnet152 layer4 (2048 units, 7x7)': (256, 2048, 7), 'layer2-like (512 units, 28x28)': (128, 512, 28),
                        'conv1-like (64 units, 112x112)': (64, 64, 112)}.items():
    acts = torch.randn(B, U, H * H, device=dev)
    top_v = torch.full((U, 15), float('-inf'), device=dev)
    top_i = torch.full((U, 15), -1, dtype=torch.long, device=dev)
    hist = torch.zeros(U, 65536, dtype=torch.int32, device=dev)
    pooled = torch.empty(B, U, device=dev)
    gb = acts.numel() * 4 / 1e9
    for label, fn in (('tally_topk', lambda: lib.milan_tally_topk(P(acts), B, U, H * H, 0, 15, P(pooled), P(top_v), P(top_i), st)),
                      ('tally_hist', lambda: lib.milan_tally_hist(P(acts), B, U, H * H, P(hist), st))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            assert fn() == 0
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name:38s} {label:11s} {gb:6.2f} GB in {ms:7.3f} ms = {gb / ms * 1e3:7.0f} GB/s')
    maps = torch.randn(2048 * 15, H, H, device=dev)
    levels = torch.zeros(len(maps), device=dev)
    masks = torch.empty(len(maps), 224, 224, dtype=torch.uint8, device=dev)
    for _ in range(2):
        lib.milan_activation_masks(P(maps), P(levels), len(maps), H, H, 224, P(masks), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.milan_activation_masks(P(maps), P(levels), len(maps), H, H, 224, P(masks), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{name:38s} masks       {masks.numel() / 1e9:6.2f} GB out in {ms:7.3f} ms = {masks.numel() / ms / 1e6:7.0f} GB/s written')
