"""Experiment: do two encoder pipelines on HALF the SMs each beat one pipeline on all of them?

The step is power-capped and alternates tensor-bound convs (3x3, reduce) with HBM-bound ones (expand + residual,
layer1 / layer2). Two engines with 74-CTA persistent kernels, fed on two streams, let an HBM-bound kernel of one chunk
run beside a tensor-bound kernel of the other. Prints images/s for: one engine on 148 SMs; two engines on 74 SMs each,
concurrently.
    python scripts/split_sm_probe.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402

sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0)
V = 5004
N = 64
images, masks = synthetic.synthetic_exemplars(N, 15, seed=5)
images, masks = images.cuda().view(-1, 3, 224, 224), masks.cuda().view(-1, 1, 224, 224)


def run(engines, streams, reps):
    for e, s in zip(engines, streams):
        with torch.cuda.stream(s):
            e.encode(images, masks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for e, s in zip(engines, streams):
            with torch.cuda.stream(s):
                e.encode(images, masks)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * len(engines) * len(images) / dt


one = Engine(sd, vocab_size=V, device='cuda:0', max_neurons=N)
print('one engine, all SMs      : %.0f images/s' % run([one], [torch.cuda.Stream()], 8))
one.close()
for sms in (74, 100):
    os.environ['MILAN_NUM_SMS'] = str(sms)
    pair = [Engine(sd, vocab_size=V, device='cuda:0', max_neurons=N) for _ in range(2)]
    print('two engines, %3d CTAs each: %.0f images/s' % (sms, run(pair, [torch.cuda.Stream(), torch.cuda.Stream()], 4)))
    print('  (one of them alone     : %.0f images/s)' % run(pair[:1], [torch.cuda.Stream()], 4))
    for e in pair:
        e.close()
