"""One warm beam + rerank decode of 64 neurons (for `ncu` launch lists): three untimed calls, then a fourth between
cudaProfilerStart / Stop so that `ncu --profile-from-start off` sees exactly one decode."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402
from oracle.make_golden import synthetic_features  # noqa: E402

feats = synthetic_features(64, 15, seed=3).cuda()
sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0, with_encoder=False)
engine = Engine(sd, vocab_size=5004, device='cuda:0', max_neurons=64)
for _ in range(3):
    out = engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
