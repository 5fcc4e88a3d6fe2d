"""Time milan_decode_beam (+rerank) on 64 neurons for logit regimes with and without early <stop> (GPU box)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402
from oracle.make_golden import VARIANTS, synthetic_features  # noqa: E402

feats = synthetic_features(64, 15, seed=3).cuda()
from neuron_descriptions_b200 import _lib  # noqa: E402

for name, fused in (('sharp', 1), ('sharp', 0), ('stop', 1), ('stop', 0), ('early', 1), ('early', 0)):
    sharpen, stop_bias = VARIANTS[name]
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
    os.environ['MILAN_FUSED_DECODE'] = str(fused)  # read at engine creation
    engine = Engine(sd, vocab_size=5004, device='cuda:0', max_neurons=64)
    for _ in range(3):
        out = engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = _lib.load().milan_launch_count()
    start.record()
    for _ in range(10):
        out = engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)
    end.record()
    torch.cuda.synchronize()
    launches = (_lib.load().milan_launch_count() - before) // 10
    print(f'{name} fused={fused}: {start.elapsed_time(end) / 10:.2f} ms per 64-neuron beam+rerank decode, '
          f'{launches} launches, group steps {out[2].tolist()}')
    engine.close()
