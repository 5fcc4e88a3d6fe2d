"""CPU: the oracle restatement (oracle/milan_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py imports /root/reference). These pin the oracle; the GPU parity tests then
compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import synthetic
from oracle import milan_oracle as O
from oracle.make_golden import VARIANTS, synthetic_features

VOCAB = synthetic.synthetic_vocab(5000)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_encoder_matches_reference_golden(golden_dir):
    g = _load(golden_dir, 'encoder_resnet101.npz')
    n, k, seed = g['meta'].tolist()
    sd = synthetic.synthetic_state_dict(seed=seed, sharpen=3.0)
    images_u8, masks_u8 = synthetic.synthetic_exemplars(n, k, seed=seed, zero_mask_fraction=0.1)
    masks_u8[0, 0] = 0
    masks_u8[1, 3, :, 100:102, 50:52] = 0
    images, masks = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        feats = O.encode(images[:1], masks[:1], sd)  # one neuron keeps the CPU suite short
    ref = torch.from_numpy(g['features'][:1])
    assert feats.shape == ref.shape == (1, k, synthetic.FEATURE_SIZE)
    # Same torch build + same weights: differences are only conv algorithm choices at different batch sizes.
    torch.testing.assert_close(feats, ref, rtol=2e-4, atol=2e-5)
    # reference test semantics (tests/milan/encoders_test.py:59-77): all-zero mask -> exactly zero features
    assert feats[0, 0].abs().max().item() == 0.0
    assert feats[0, 1].abs().max().item() > 0.0


@pytest.mark.parametrize('kind,arch', [('pyramid', 'resnet18'), ('pyramid', 'resnet50'), ('spatial', 'resnet18'),
                                       ('pyramid', 'alexnet')])
def test_encoder_variants_match_reference_golden(golden_dir, kind, arch):
    """The secondary encoder configs (`src/milan/encoders.py:214-216,326-351`) of the oracle vs the reference."""
    from oracle.make_golden import encoder_variant_inputs
    g = _load(golden_dir, 'encoder_variants.npz')
    images_u8, masks_u8 = encoder_variant_inputs()
    images, masks = O.to_float_inputs(images_u8, masks_u8)
    sd = {'encoder.' + k: v for k, v in synthetic.synthetic_encoder_state_dict(arch, seed=3).items()}
    with torch.no_grad():
        if kind == 'spatial':
            feats = O.spatial_encode(images, masks, sd, arch)
        else:
            feats = O.pyramid_encode(images, masks, sd, arch)
    ref = torch.from_numpy(g[f'{kind}_{arch}'])
    torch.testing.assert_close(feats, ref, rtol=2e-4, atol=2e-5)
    if kind == 'pyramid':
        assert feats[1].abs().max().item() == 0.0  # all-zero mask
        assert feats[2].abs().max().item() > 0.0


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_decoder_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, f'decoder_{name}.npz')
    n, k, stop_index = g['meta'].tolist()
    sharpen, stop_bias = VARIANTS[name]
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
    feats = synthetic_features(n, k, seed=0)

    state = O.init_state(feats, sd)
    np.testing.assert_allclose(state.h.numpy(), g['init_h'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(state.c.numpy(), g['init_c'], rtol=1e-5, atol=1e-6)
    first = O.step(feats, torch.full((n,), len(VOCAB), dtype=torch.long), state, sd)
    np.testing.assert_allclose(first.attentions.numpy(), g['step0_attn'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(first.state.h.numpy(), g['step0_h'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(first.predictions.logsumexp(-1).numpy(), g['step0_logsumexp'], atol=1e-5)
    top_v, top_i = first.predictions.topk(8, dim=-1)
    np.testing.assert_array_equal(top_i.numpy(), g['step0_top_indices'])
    np.testing.assert_allclose(top_v.numpy(), g['step0_top_values'], atol=1e-5)

    greedy = O.decode(feats, sd, VOCAB, strategy='greedy', mi=False)
    np.testing.assert_array_equal(greedy.tokens.numpy(), g['greedy_tokens'])
    np.testing.assert_allclose(greedy.scores.numpy(), g['greedy_scores'], atol=1e-4)
    np.testing.assert_allclose(greedy.attentions.numpy(), g['greedy_attn'], rtol=1e-4, atol=1e-6)
    assert list(greedy.captions) == list(g['greedy_captions'])

    greedy_mi = O.decode(feats, sd, VOCAB, strategy='greedy', mi=True)
    np.testing.assert_array_equal(greedy_mi.tokens.numpy(), g['greedy_mi_tokens'])
    np.testing.assert_allclose(greedy_mi.scores.numpy(), g['greedy_mi_scores'], atol=1e-4)

    rerank = O.decode(feats, sd, VOCAB, strategy='rerank', beam_size=50)
    np.testing.assert_array_equal(rerank.beam_tokens.numpy(), g['beam_tokens'])
    np.testing.assert_allclose(rerank.beam_scores.numpy(), g['beam_scores'], atol=1e-4)
    np.testing.assert_array_equal(rerank.tokens.numpy(), g['rerank_tokens'])
    np.testing.assert_allclose(rerank.scores.numpy(), g['rerank_scores'], atol=1e-4)
    assert list(rerank.captions) == list(g['captions'])

    inputs_lm = torch.cat([torch.full((n * 50, 1), len(VOCAB), dtype=torch.long),
                           rerank.beam_tokens.view(n * 50, -1)], dim=-1)
    np.testing.assert_allclose(O.lm_forward(inputs_lm, sd, stop_index).numpy(), g['lm_scores'], atol=1e-4)

    beam_mi = O.decode(feats, sd, VOCAB, strategy='beam', beam_size=10)
    np.testing.assert_array_equal(beam_mi.beam_tokens.numpy(), g['beam_mi_tokens'])
    np.testing.assert_allclose(beam_mi.beam_scores.numpy(), g['beam_mi_scores'], atol=1e-4)

    small = O.decode(feats, sd, VOCAB, strategy='rerank', beam_size=7, length=9)
    np.testing.assert_array_equal(small.beam_tokens.numpy(), g['small_beam_tokens'])
    np.testing.assert_array_equal(small.tokens.numpy(), g['small_rerank_tokens'])


def test_beam_search_indirect_pins():
    """allennlp is unpinned offline; check the invariants its algorithm guarantees (oracle/beam_search.py header)."""
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=2.0, with_encoder=False)
    feats = synthetic_features(3, 15, seed=1)
    out = O.decode(feats, sd, VOCAB, strategy='beam', beam_size=5, length=8, mi=False)
    bt, bs = out.beam_tokens, out.beam_scores
    assert (bs[:, :-1] >= bs[:, 1:]).all()
    # each beam score == forced-decode log-prob sum of its sequence (stop-padded tails cost 0)
    stop = len(VOCAB) + 1
    for b in range(bt.shape[1]):
        forced = O.decode(feats, sd, VOCAB, strategy=bt[:, b], length=bt.shape[-1], mi=False)
        chosen = forced.predictions.gather(2, bt[:, b].unsqueeze(-1)).squeeze(-1)
        ended = (bt[:, b] == stop).long().cumsum(-1) - (bt[:, b] == stop).long() > 0  # strictly after first stop
        total = chosen.masked_fill(ended, 0.0).sum(-1)
        torch.testing.assert_close(total, bs[:, b], atol=2e-4, rtol=0)
    # beam_size=1 == greedy up to the first <stop>
    one = O.decode(feats, sd, VOCAB, strategy='beam', beam_size=1, length=8, mi=False)
    greedy = O.decode(feats, sd, VOCAB, strategy='greedy', length=8, mi=False)
    for row_b, row_g in zip(one.beam_tokens[:, 0].tolist(), greedy.tokens.tolist()):
        cut = row_g.index(stop) + 1 if stop in row_g else len(row_g)
        assert row_b[:cut] == row_g[:cut]


def test_reconstruct_matches_reference_rules():
    vocab = ('.', ',', '-', 'dog', 'cat', 'top')
    n = len(vocab)
    assert O.reconstruct([3, 1, 4, 0, 5, 2, 3, n + 1, 4], vocab) == 'Dog, cat. Top-dog'
    assert O.reconstruct([n, 3, n + 2, n + 3, 4], vocab) == 'Dog cat'
    with pytest.raises(ValueError):
        O.reconstruct([n + 9], vocab)


@pytest.mark.parametrize('layer,output_size,k', [('conv_1', 16, 4), ('conv_2', 24, 3)])
def test_exemplars_oracle_matches_reference_golden(golden_dir, layer, output_size, k):
    """Stage-1 restatement (`oracle/exemplars_oracle.py`) vs the reference's `exemplars.compute.discriminative`."""
    from oracle import exemplars_oracle as E
    from oracle.make_golden import exemplar_toy_images, exemplar_toy_model
    g = _load(golden_dir, 'exemplars.npz')
    model, images = exemplar_toy_model(), exemplar_toy_images()
    features = (lambda x: model.conv_1(x)) if layer == 'conv_1' else (lambda x: model(x))
    out = E.discriminative(features, images, k, 0.99, output_size, batch_size=8)
    np.testing.assert_array_equal(out['ids'], g[f'{layer}_ids'])
    np.testing.assert_allclose(out['activations'], g[f'{layer}_activations'], rtol=2e-5, atol=2e-6)  # csv: %.5e
    np.testing.assert_array_equal(out['images'], g[f'{layer}_images'])
    np.testing.assert_array_equal(out['masks'], g[f'{layer}_masks'])


@pytest.mark.parametrize('layer,output_size,k,units', [('conv_2', 24, 3, None), ('conv_1', 16, 4, (0, 2, 5))])
def test_generative_exemplars_oracle_matches_reference_golden(golden_dir, layer, output_size, k, units):
    """`exemplars.compute.generative` (`src/exemplars/compute.py:352-437`): kept images are the generator's outputs."""
    from oracle import exemplars_oracle as E
    from oracle.make_golden import exemplar_toy_images, generative_toy_model
    g = _load(golden_dir, 'exemplars_generative.npz')
    model, zs = generative_toy_model(), exemplar_toy_images(seed=11)
    features = (lambda x: model.conv_1(x)) if layer == 'conv_1' else (lambda x: model.conv_2(model.conv_1(x)))
    with torch.no_grad():
        out = E.discriminative(features, zs, k, 0.99, output_size, batch_size=8, images_fn=model, units=units)
    np.testing.assert_array_equal(out['ids'], g[f'{layer}_ids'])
    np.testing.assert_allclose(out['activations'], g[f'{layer}_activations'], rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(out['images'], g[f'{layer}_images'])
    np.testing.assert_array_equal(out['masks'], g[f'{layer}_masks'])
    if units is not None:
        np.testing.assert_array_equal(g[f'{layer}_units'], sorted(units))


def test_vit_exemplars_oracle_matches_reference_golden(golden_dir):
    """`discriminative(transform_hiddens=spatialize_vit_mlp)`: the DINO ViT-S/8 configuration of
    `src/exemplars/models.py:236-247` on a toy ViT (CLS token dropped, patches arranged on their grid)."""
    from neuron_descriptions_b200.exemplars import transforms
    from oracle import exemplars_oracle as E
    from oracle.make_golden import VIT_CASE, ToyViT, exemplar_toy_images
    g = _load(golden_dir, 'exemplars_vit.npz')
    model, images = ToyViT().eval(), exemplar_toy_images(seed=23)
    _, output_size, k = VIT_CASE
    retained = {}
    model.mlp.fc1.register_forward_hook(lambda _m, _i, output: retained.__setitem__('x', output))

    def features(x):
        model(x)
        return transforms.spatialize_vit_mlp(retained['x'])

    with torch.no_grad():
        out = E.discriminative(features, images, k, 0.99, output_size, batch_size=8)
    assert out['masks'].shape == (10, k, 1, output_size, output_size)  # 10 MLP units on the 4 x 4 patch grid
    np.testing.assert_array_equal(out['ids'], g['ids'])
    np.testing.assert_allclose(out['activations'], g['activations'], rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(out['images'], g['images'])
    np.testing.assert_array_equal(out['masks'], g['masks'])
