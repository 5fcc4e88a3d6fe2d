"""GPU: the fused beam step / LM rerank (csrc/decode_fused.cu, EPI_LSTM / EPI_HEAD epilogues) against the
one-kernel-per-op path it replaces (`MILAN_FUSED_DECODE=0`, read when an engine is created) and the oracle.

The reference-golden comparisons of tests/test_gpu_parity.py run through the fused path by default; these cases
add the A/B against the unfused kernels and the inputs that drive the exact-selection fallback (mass ties)."""
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import synthetic
from oracle import milan_oracle as O
from oracle.make_golden import VARIANTS, synthetic_features
from test_gpu_parity import _tokens_match  # tests/ is on sys.path (pytest rootdir import mode)

pytestmark = pytest.mark.gpu

VOCAB = synthetic.synthetic_vocab(5000)
V = len(VOCAB) + 4
STOP = len(VOCAB) + 1


def _engine(sd, fused, **kwargs):
    from neuron_descriptions_b200.engine import Engine
    old = os.environ.get('MILAN_FUSED_DECODE')
    os.environ['MILAN_FUSED_DECODE'] = '1' if fused else '0'
    try:
        return Engine(sd, vocab_size=V, device='cuda:0', **kwargs)
    finally:
        if old is None:
            del os.environ['MILAN_FUSED_DECODE']
        else:
            os.environ['MILAN_FUSED_DECODE'] = old


@pytest.mark.parametrize('name', ['sharp', 'stop', 'early'])
def test_fused_beam_matches_unfused(name):
    sharpen, stop_bias = VARIANTS[name]
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
    feats = synthetic_features(24, 15, seed=4)
    fused, plain = _engine(sd, True, max_neurons=24), _engine(sd, False, max_neurons=24)
    for beam, length, group in ((50, 15, 16), (7, 9, 24), (1, 15, 8), (33, 1, 16)):
        a = fused.decode_beam(feats, length, beam, True, 0.2, group_size=group)
        b = plain.decode_beam(feats, length, beam, True, 0.2, group_size=group)
        assert torch.equal(a[2], b[2]), 'early-exit lengths differ'
        # the two paths round differently (fast tanh / sigmoid, softmax from partials): sequences may swap only
        # where their scores tie, exactly the allowance the reference-golden tests make
        torch.testing.assert_close(a[1], b[1], atol=1e-4, rtol=0)
        _tokens_match(a[0].cpu().numpy(), b[0].cpu().numpy(), a[1].cpu().numpy(), b[1].cpu().numpy(), f'beam {beam}')
        torch.testing.assert_close(a[4], b[4], atol=1e-4, rtol=0)
        _tokens_match(a[3].cpu().numpy(), b[3].cpu().numpy(), a[4].cpu().numpy(), b[4].cpu().numpy(), f'rerank {beam}')
        same = (a[0] == b[0]).all(-1)
        torch.testing.assert_close(a[5][same], b[5][same], atol=1e-4, rtol=0)  # LM scores of identical sequences
        print(f'{name} beam {beam}: {int((~same).sum())} of {same.numel()} beam sequences differ (score ties)')
    fused.close()
    plain.close()


def test_fused_lm_score_matches_unfused_and_oracle():
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=1.0, with_encoder=False)
    gen = torch.Generator().manual_seed(7)
    inputs = torch.randint(0, len(VOCAB), (300, 16), generator=gen)  # more rows than one 128-row tile
    inputs[:, 0] = len(VOCAB)
    inputs[3, 4] = STOP
    inputs[5, 1] = STOP
    inputs[7, 15] = STOP
    inputs[9, 5:] = STOP
    inputs[200:, 8:] = STOP
    fused, plain = _engine(sd, True, max_neurons=8), _engine(sd, False, max_neurons=8)
    a, b = fused.lm_score(inputs).cpu(), plain.lm_score(inputs).cpu()
    torch.testing.assert_close(a, b, atol=1e-4, rtol=0)
    torch.testing.assert_close(a, O.lm_forward(inputs, sd, STOP), atol=1e-3, rtol=0)
    fused.close()
    plain.close()


def test_fused_exact_selection_under_mass_ties():
    """Logits with only a handful of distinct values: thousands of classes tie at the 50th place, the prefilter of
    beam_select overflows and the radix-select fallback must reproduce the unfused kernel's choice (ties -> lower
    class index) exactly."""
    sd = synthetic.synthetic_state_dict(seed=3, sharpen=1.0, stop_bias=0.0, with_encoder=False)
    gen = torch.Generator().manual_seed(11)
    sd['output.1.weight'] = torch.zeros_like(sd['output.1.weight'])
    sd['output.1.bias'] = torch.randint(0, 4, sd['output.1.bias'].shape, generator=gen).float()
    feats = synthetic_features(5, 15, seed=2)
    fused, plain = _engine(sd, True, max_neurons=8), _engine(sd, False, max_neurons=8)
    for beam in (50, 3):
        a = fused.decode_beam(feats, 6, beam, False, 0.2)
        b = plain.decode_beam(feats, 6, beam, False, 0.2)
        assert torch.equal(a[0], b[0])
        torch.testing.assert_close(a[1], b[1], atol=1e-5, rtol=0)
    # first step: the beam is the lowest-index members of the top bias level (every log-probability there ties)
    first = fused.decode_beam(feats, 1, 50, False, 0.2)[0][0, :, 0].cpu()
    assert torch.equal(first, (sd['output.1.bias'] == 3).nonzero().flatten()[:50])
    fused.close()
    plain.close()


def test_fused_step_launch_count():
    """North-star shape of the step: four launches per beam step, three per LM position."""
    from neuron_descriptions_b200 import _lib
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=3.0, with_encoder=False)
    feats = synthetic_features(16, 15, seed=1)
    engine = _engine(sd, True, max_neurons=16)
    lib = _lib.load()
    engine.decode_beam(feats, 15, 50, True, 0.2)
    before = lib.milan_launch_count()
    engine.decode_beam(feats, 15, 50, True, 0.2)
    torch.cuda.synchronize()
    launches = lib.milan_launch_count() - before
    # prepare (split, kh GEMM, mean, init GEMM, init finish, fill, q/g GEMM) + 15 x 4 + backtrack x 2
    # + lm_skip + 15 x 3 + finalize + select
    assert launches <= 7 + 15 * 4 + 2 + 1 + 15 * 3 + 2, launches
    engine.close()
