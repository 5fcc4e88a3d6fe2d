"""Tiny end-to-end workload for `compute-sanitizer` (memcheck / racecheck): every kernel family once.

  compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402

vocab = synthetic.synthetic_vocab(300)
V = len(vocab) + 4
# flagship: resnet101 pyramid + beam/rerank through the pipelined host call (3 chunks of 1 neuron)
sd = synthetic.synthetic_state_dict(seed=0, vocab_size=300, sharpen=12.0, stop_bias=1.0)
engine = Engine(sd, vocab_size=V, device='cuda:0', max_neurons=1, max_beam=4, max_keys=2, max_length=6)
images, masks = synthetic.synthetic_exemplars(3, 2, seed=1)
tokens, scores, steps = engine.describe_host(images, masks, strategy='rerank', beam=4, length=6, group_size=1)
print('describe_host tokens', tokens[0].tolist(), 'scores', scores.tolist(), 'steps', steps.tolist())
tokens, scores, _ = engine.describe_host(images, masks, strategy='greedy', mi=True, length=6)
print('greedy-mi scores', scores.tolist())
engine.close()
# secondary encoders (encoder-only engines)
for kind, arch, F in (('pyramid', 'resnet18', 1024), ('spatial', 'resnet18', 512), ('pyramid', 'alexnet', 1152)):
    esd = {'encoder.' + k: v for k, v in synthetic.synthetic_encoder_state_dict(arch, seed=3).items()}
    enc = Engine(esd, vocab_size=68, device='cuda:0', feature_size=F, encoder_arch=arch, encoder_kind=kind,
                 max_neurons=1, max_beam=1, max_keys=1, max_length=1, max_images=2, decoder=False)
    feats = enc.encode(images[0, :2], masks[0, :2])
    print(kind, arch, tuple(feats.shape), float(feats.abs().max()))
    enc.close()
torch.cuda.synchronize()
print('done')
