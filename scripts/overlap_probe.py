"""Probe: how much of the decode phase hides under the next batch's encoder when they run on two streams?

  python scripts/overlap_probe.py [--neurons 64] [--iters 6]

Prints ms per batch for: encode only, decode only, sequential (one stream), overlapped (decode of batch i on a
high-priority stream while batch i+1 is encoded). Measurement aid only (not a bench number).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--neurons', type=int, default=64)
    parser.add_argument('--iters', type=int, default=6)
    args = parser.parse_args()
    nb, iters = args.neurons, args.iters
    device = torch.device('cuda', 0)
    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0)
    engine = Engine(sd, vocab_size=len(vocab) + 4, device=device, max_neurons=nb)
    batches = []
    for i in range(2):
        images, masks = synthetic.synthetic_exemplars(nb, 15, seed=i)
        batches.append((images.to(device).view(-1, 3, 224, 224), masks.to(device).view(-1, 1, 224, 224)))

    def encode(i):
        images, masks = batches[i % 2]
        return engine.encode(images, masks).view(nb, 15, -1)

    def decode(feats):
        return engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)[3]

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / iters

    feats0 = encode(0)
    torch.cuda.synchronize()
    out = {}
    out['encode_ms'] = timed(lambda: [encode(i) for i in range(iters)])
    out['decode_ms'] = timed(lambda: [decode(feats0) for _ in range(iters)])
    out['sequential_ms'] = timed(lambda: [decode(encode(i)) for i in range(iters)])

    hi = torch.cuda.Stream(device, priority=-1)
    main_stream = torch.cuda.current_stream(device)

    def overlapped():
        feats = encode(0)
        for i in range(1, iters + 1):
            ready = torch.cuda.Event()
            ready.record(main_stream)
            with torch.cuda.stream(hi):
                hi.wait_event(ready)
                feats.record_stream(hi)
                decode(feats)
            if i < iters:
                feats = encode(i)
        main_stream.wait_stream(hi)

    out['overlapped_ms'] = timed(overlapped)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
