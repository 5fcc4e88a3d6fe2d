"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference-generated goldens.

Tolerances (north_star): identical greedy token ids; beam / greedy log-prob sums within 1e-3 (fp32);
encoder features within 1e-3 relative to the feature scale (the convs run as bf16 hi/lo split products with
fp32 accumulation; measured error is ~1e-5).
"""
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import synthetic
from oracle import milan_oracle as O
from oracle.make_golden import VARIANTS, synthetic_features

pytestmark = pytest.mark.gpu

VOCAB = synthetic.synthetic_vocab(5000)
V = len(VOCAB) + 4
STOP = len(VOCAB) + 1
LOGP_TOL = 1e-3


def _engine(sd, **kwargs):
    from neuron_descriptions_b200.engine import Engine
    return Engine(sd, vocab_size=V, device='cuda:0', **kwargs)


@pytest.fixture(scope='module')
def full_sd():
    return synthetic.synthetic_state_dict(seed=0, sharpen=3.0)


@pytest.fixture(scope='module')
def full_engine(full_sd):
    engine = _engine(full_sd, max_neurons=16)
    yield engine
    engine.close()


def _golden_inputs():
    images_u8, masks_u8 = synthetic.synthetic_exemplars(2, 15, seed=0, zero_mask_fraction=0.1)
    masks_u8[0, 0] = 0
    masks_u8[1, 3, :, 100:102, 50:52] = 0
    return images_u8, masks_u8


def test_encoder_matches_reference_golden(full_engine, golden_dir):
    g = np.load(os.path.join(golden_dir, 'encoder_resnet101.npz'))
    images_u8, masks_u8 = _golden_inputs()
    feats = full_engine.encode(images_u8.view(-1, 3, 224, 224), masks_u8.view(-1, 1, 224, 224)).cpu().view(2, 15, -1)
    ref = torch.from_numpy(g['features'])
    scale = ref.abs().max().item()
    err = (feats - ref).abs().max().item()
    print(f'encoder max abs err {err:.3e} (feature scale {scale:.3f})')
    assert err <= 1e-3 * scale, f'encoder features differ from the reference golden: {err} (scale {scale})'
    torch.testing.assert_close(feats, ref, rtol=2e-3, atol=1e-3 * scale)
    # all-zero mask -> exactly zero features (reference tests/milan/encoders_test.py:59-77)
    assert feats[0, 0].abs().max().item() == 0.0
    assert feats[0, 1].abs().max().item() > 0.0


def test_encoder_float_and_uint8_inputs_agree(full_engine):
    images_u8, masks_u8 = _golden_inputs()
    images_u8, masks_u8 = images_u8[:1].reshape(-1, 3, 224, 224), masks_u8[:1].reshape(-1, 1, 224, 224)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    a = full_engine.encode(images_u8, masks_u8)
    b = full_engine.encode(images_f, masks_f)
    assert torch.equal(a, b), (a - b).abs().max()
    # masks=None == all-ones masks (src/milan/encoders.py:292-293)
    c = full_engine.encode(images_u8[:3], None)
    d = full_engine.encode(images_u8[:3], torch.ones_like(masks_u8[:3]))
    assert torch.equal(c, d)


def test_encoder_matches_oracle_on_fresh_inputs(full_engine, full_sd):
    images_u8, masks_u8 = synthetic.synthetic_exemplars(1, 6, seed=5)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        ref = O.encode(images_f, masks_f, full_sd)
    got = full_engine.encode(images_u8.view(-1, 3, 224, 224), masks_u8.view(-1, 1, 224, 224)).cpu().view(1, 6, -1)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 1e-3 * scale


@pytest.fixture(scope='module', params=sorted(VARIANTS))
def variant(request, golden_dir):
    name = request.param
    sharpen, stop_bias = VARIANTS[name]
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
    engine = _engine(sd, max_neurons=16)
    g = np.load(os.path.join(golden_dir, f'decoder_{name}.npz'))
    yield name, sd, engine, g
    engine.close()


def _tokens_match(got, ref, got_scores, ref_scores, what):
    """Token ids must be identical except where the reference itself is within rounding of a tie."""
    got, ref = np.asarray(got), np.asarray(ref)
    if np.array_equal(got, ref):
        return
    bad = np.argwhere((got != ref).any(axis=-1))
    for idx in bad:
        idx = tuple(idx)
        assert abs(float(got_scores[idx]) - float(ref_scores[idx])) <= LOGP_TOL, (
            f'{what}: sequence {idx} differs and scores are not tied: {got[idx]} vs {ref[idx]} '
            f'({got_scores[idx]} vs {ref_scores[idx]})')
    frac = len(bad) / max(1, int(np.prod(got.shape[:-1])))
    assert frac <= 0.02, f'{what}: {len(bad)} sequences differ (near-ties), too many'


def test_decoder_init_and_first_step(variant):
    name, sd, engine, g = variant
    n, k, _ = g['meta'].tolist()
    feats = synthetic_features(n, k, seed=0)
    h, c = engine.init_state(feats)
    np.testing.assert_allclose(h.cpu().numpy(), g['init_h'], atol=2e-5)
    np.testing.assert_allclose(c.cpu().numpy(), g['init_c'], atol=2e-5)
    start = torch.full((n,), len(VOCAB), dtype=torch.long)
    pred, attn, h1, c1, _, _ = engine.step(feats, start, h, c)
    np.testing.assert_allclose(attn.cpu().numpy(), g['step0_attn'], atol=2e-5)
    np.testing.assert_allclose(h1.cpu().numpy(), g['step0_h'], atol=2e-5)
    np.testing.assert_allclose(c1.cpu().numpy(), g['step0_c'], atol=5e-5)
    np.testing.assert_allclose(pred.logsumexp(-1).cpu().numpy(), g['step0_logsumexp'], atol=1e-4)
    top_v, top_i = pred.cpu().topk(8, dim=-1)
    np.testing.assert_allclose(top_v.numpy(), g['step0_top_values'], atol=2e-4)
    if name != 'flat':
        np.testing.assert_array_equal(top_i.numpy(), g['step0_top_indices'])


def test_decoder_greedy(variant):
    name, sd, engine, g = variant
    n, k, _ = g['meta'].tolist()
    feats = synthetic_features(n, k, seed=0)
    tokens, scores, predictions, attentions = engine.decode_greedy(feats, 15, mi=False, temperature=0.2)
    if name == 'flat':  # near-uniform logits: argmax is rounding-sensitive, scores are still pinned
        _tokens_match(tokens.cpu().numpy(), g['greedy_tokens'], scores.cpu().numpy(), g['greedy_scores'], 'greedy')
    else:
        np.testing.assert_array_equal(tokens.cpu().numpy(), g['greedy_tokens'])
        np.testing.assert_allclose(scores.cpu().numpy(), g['greedy_scores'], atol=LOGP_TOL)
        np.testing.assert_allclose(attentions.cpu().numpy(), g['greedy_attn'], atol=1e-4)
        chosen = predictions.gather(2, tokens.unsqueeze(-1)).squeeze(-1)
        np.testing.assert_allclose(chosen.cpu().numpy(), g['greedy_chosen_logp'], atol=LOGP_TOL)
    # MI greedy (the reference default for strategy='greedy' when the decoder has an LM, decoders.py:385-387)
    tokens_mi, scores_mi, _, _ = engine.decode_greedy(feats, 15, mi=True, temperature=0.2)
    if name != 'flat':
        np.testing.assert_array_equal(tokens_mi.cpu().numpy(), g['greedy_mi_tokens'])
        np.testing.assert_allclose(scores_mi.cpu().numpy(), g['greedy_mi_scores'], atol=LOGP_TOL)


def test_decoder_beam_and_rerank(variant):
    name, sd, engine, g = variant
    n, k, _ = g['meta'].tolist()
    feats = synthetic_features(n, k, seed=0)
    beam_tokens, beam_scores, steps, tokens, scores, lm_scores = engine.decode_beam(feats, 15, 50, True, 0.2)
    T = int(steps[0].item())
    assert T == g['beam_tokens'].shape[-1], f'early-exit length {T} vs reference {g["beam_tokens"].shape[-1]}'
    np.testing.assert_allclose(beam_scores.cpu().numpy(), g['beam_scores'], atol=LOGP_TOL)
    if name != 'flat':
        _tokens_match(beam_tokens[..., :T].cpu().numpy(), g['beam_tokens'], beam_scores.cpu().numpy(), g['beam_scores'],
                      'beam')
        np.testing.assert_allclose(lm_scores.view(-1).cpu().numpy(), g['lm_scores'], atol=LOGP_TOL)
        np.testing.assert_array_equal(tokens[..., :T].cpu().numpy(), g['rerank_tokens'])
    np.testing.assert_allclose(scores.cpu().numpy(), g['rerank_scores'], atol=LOGP_TOL)
    assert (tokens[..., T:] == STOP).all()
    # smaller beam / shorter length
    bt, bs, steps, tok, sc, _ = engine.decode_beam(feats, 9, 7, True, 0.2)
    T = int(steps[0].item())
    assert T == g['small_beam_tokens'].shape[-1]
    np.testing.assert_allclose(bs.cpu().numpy(), g['small_beam_scores'], atol=LOGP_TOL)
    np.testing.assert_allclose(sc.cpu().numpy(), g['small_rerank_scores'], atol=LOGP_TOL)
    if name != 'flat':
        np.testing.assert_array_equal(bt[..., :T].cpu().numpy(), g['small_beam_tokens'])
        np.testing.assert_array_equal(tok[..., :T].cpu().numpy(), g['small_rerank_tokens'])


def test_decoder_mi_beam(variant):
    """strategy='beam' with an LM defaults to MI decoding: the LM state follows the beam (decoders.py:385-387)."""
    name, sd, engine, g = variant
    n, k, _ = g['meta'].tolist()
    feats = synthetic_features(n, k, seed=0)
    bt, bs, steps, tok, sc, _ = engine.decode_beam(feats, 15, 10, False, 0.2, mi=True)
    T = int(steps[0].item())
    assert T == g['beam_mi_tokens'].shape[-1]
    np.testing.assert_allclose(bs.cpu().numpy(), g['beam_mi_scores'], atol=LOGP_TOL)
    if name != 'flat':
        _tokens_match(bt[..., :T].cpu().numpy(), g['beam_mi_tokens'], bs.cpu().numpy(), g['beam_mi_scores'], 'mi-beam')
        np.testing.assert_array_equal(tok[..., :T].cpu().numpy(), g['beam_mi_tokens'][:, 0])
    with pytest.raises(ValueError, match='cannot set `mi=` decoding when reranking'):
        engine.decode_beam(feats, 15, 10, True, 0.2, mi=True)


def test_lm_score_matches_oracle(variant):
    name, sd, engine, g = variant
    gen = torch.Generator().manual_seed(3)
    inputs = torch.randint(0, len(VOCAB), (37, 12), generator=gen)
    inputs[:, 0] = len(VOCAB)
    inputs[3, 4] = STOP
    inputs[5, 1] = STOP
    inputs[7, 11] = STOP
    inputs[9, 5:] = STOP
    ref = O.lm_forward(inputs, sd, STOP)
    got = engine.lm_score(inputs).cpu()
    torch.testing.assert_close(got, ref, atol=LOGP_TOL, rtol=0)


def test_lm_logprobs_and_custom_masks_match_oracle(variant):
    """`LanguageModel.forward(reduce=False)` and `forward(reduce=True, masks=...)` (`src/milan/lms.py:58-101`)."""
    from neuron_descriptions_b200.milan import lms
    name, sd, engine, g = variant
    gen = torch.Generator().manual_seed(5)
    inputs = torch.randint(0, len(VOCAB), (9, 7), generator=gen)
    inputs[:, 0] = len(VOCAB)
    inputs[2, 3] = STOP
    ref = O.lm_forward(inputs, sd, STOP, reduce=False)
    lm = lms.LanguageModel(None).bind(engine)
    got = lm(inputs).cpu()
    assert got.shape == ref.shape == (9, 7, len(VOCAB) + 4)
    torch.testing.assert_close(got, ref, atol=2e-4, rtol=0)
    masks = (torch.rand(9, 6, generator=gen) > 0.3).long()
    want = ref[:, :-1].gather(2, inputs[:, 1:].unsqueeze(-1)).squeeze(-1).mul(masks).sum(-1)
    torch.testing.assert_close(lm(inputs, reduce=True, masks=masks).cpu(), want, atol=LOGP_TOL, rtol=0)
    torch.testing.assert_close(lm(inputs, reduce=True).cpu(), O.lm_forward(inputs, sd, STOP), atol=LOGP_TOL, rtol=0)


def test_decoder_with_reranker_contract():
    """`DecoderWithCLIP.forward` (`src/milan/decoders.py:1135-1196`) with an injected similarity model: captions, scores
    and tokens are those of the beam entry the reranker ranks first; everything else is the beam decode's."""
    from neuron_descriptions_b200 import milan
    from neuron_descriptions_b200.milan import lang, rerankers
    from oracle.make_golden import reranker_similarity
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=1.0)
    indexer = lang.Indexer(lang.Vocab(VOCAB), start=True, stop=True, pad=True, unk=True)
    base = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                         lm=milan.LanguageModel(indexer), max_neurons=4)
    base.load_state_dict(sd)
    reranker = rerankers.SimilarityReranker(reranker_similarity, lam=.4)
    decoder = milan.DecoderWithCLIP.from_decoder(base, reranker=reranker, max_neurons=4)
    assert decoder.reranker is reranker and decoder.beam_size == base.beam_size  # properties travel with the payload
    decoder.to('cuda:0')
    images_u8, masks_u8 = synthetic.synthetic_exemplars(3, 15, seed=8)
    images, masks = O.to_float_inputs(images_u8, masks_u8)
    with pytest.raises(ValueError, match='must specify masks'):
        decoder(images)
    with pytest.raises(ValueError, match='cannot set "strategy"'):
        decoder(images, masks, strategy='greedy')
    out = decoder(images, masks, beam_size=12, mi=False)
    beam = milan.Decoder.forward(decoder, images, masks=masks, strategy='beam', beam_size=12, mi=False)
    ranked = reranker(images, masks, beam.beam_captions)
    for i in range(3):
        first = ranked.orders[i][0]
        assert out.captions[i] == beam.beam_captions[i][first] == ranked.texts[i][0]
        assert torch.equal(out.tokens[i], beam.beam_tokens[i, first])
        assert float(out.scores[i]) == float(beam.beam_scores[i, first])
    assert torch.equal(out.beam_tokens, beam.beam_tokens) and len(out.beam_captions) == 3
    assert 'reranker_kwargs' in decoder.properties()
    with pytest.raises(NotImplementedError, match='clip'):
        milan.DecoderWithCLIP(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False))


def test_beam_properties_full_size():
    """Size-independent properties at BASELINE size (16 neurons x beam 50 x 15 steps, V = 5004)."""
    sd = synthetic.synthetic_state_dict(seed=1, sharpen=12.0, stop_bias=2.0, with_encoder=False)
    engine = _engine(sd, max_neurons=32)
    feats = synthetic_features(32, 15, seed=9)
    bt, bs, steps, tok, sc, lm = engine.decode_beam(feats, 15, 50, True, 0.2, group_size=16)
    assert (bs[:, :-1] >= bs[:, 1:]).all(), 'beam scores must be sorted descending'
    # every beam score equals the forced-decode log-prob sum of its sequence (tokens after the first <stop> cost 0)
    for j in (0, 17, 49):
        seq = bt[:, j]
        _, _, pred, _ = engine.decode_greedy(feats, 15, mi=False, temperature=0.2, forced=seq)
        chosen = pred.gather(2, seq.unsqueeze(-1)).squeeze(-1)
        after = ((seq == STOP).long().cumsum(-1) - (seq == STOP).long()) > 0
        total = chosen.masked_fill(after, 0.0).sum(-1)
        torch.testing.assert_close(total, bs[:, j], atol=LOGP_TOL, rtol=0)
    # rerank score = beam score - T * lm score at the chosen index (decoders.py:507)
    combined = bs - 0.2 * lm
    torch.testing.assert_close(sc, combined.max(dim=1).values, atol=1e-5, rtol=0)
    # grouping: decoding the second group alone gives the same sequences (neurons are independent)
    bt2, bs2, _, _, _, _ = engine.decode_beam(feats[16:], 15, 50, False, 0.2)
    assert torch.equal(bt2, bt[16:]) and torch.allclose(bs2, bs[16:], atol=1e-5)
    # beam_size = 1 reproduces greedy up to the first <stop>
    b1, _, _, _, _, _ = engine.decode_beam(feats, 15, 1, False, 0.2)
    gt, _, _, _ = engine.decode_greedy(feats, 15, mi=False, temperature=0.2)
    for row_b, row_g in zip(b1[:, 0].tolist(), gt.tolist()):
        cut = row_g.index(STOP) + 1 if STOP in row_g else len(row_g)
        assert row_b[:cut] == row_g[:cut]
    engine.close()


def test_describe_host_end_to_end(full_engine, full_sd):
    """milan_describe_host (host uint8 in, token ids out) vs the oracle's predict on the same exemplars."""
    images_u8, masks_u8 = synthetic.synthetic_exemplars(3, 15, seed=11)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        feats = O.encode(images_f, masks_f, full_sd)
        ref = O.decode(feats, full_sd, VOCAB, strategy='rerank', beam_size=50)
        ref_greedy = O.decode(feats, full_sd, VOCAB, strategy='greedy', mi=False)
    tokens, scores, steps = full_engine.describe_host(images_u8, masks_u8, strategy='rerank')
    T = ref.tokens.shape[-1]
    assert int(steps[0]) == T
    torch.testing.assert_close(scores, ref.scores, atol=LOGP_TOL, rtol=0)
    _tokens_match(tokens[:, :T].numpy(), ref.tokens.numpy(), scores.numpy(), ref.scores.numpy(), 'describe/rerank')
    tokens_g, scores_g, _ = full_engine.describe_host(images_u8, masks_u8, strategy='greedy', mi=False)
    torch.testing.assert_close(scores_g, ref_greedy.scores, atol=LOGP_TOL, rtol=0)
    _tokens_match(tokens_g.numpy(), ref_greedy.tokens.numpy(), scores_g.numpy(), ref_greedy.scores.numpy(),
                  'describe/greedy')


def test_describe_host_pipelined_chunks(full_sd):
    """Several chunks through the double-buffered copy/compute pipeline of milan_describe_host give exactly the
    results of the same chunks described one call at a time (ragged last chunk, pinned and pageable inputs)."""
    engine = _engine(full_sd, max_neurons=4, max_beam=8, max_keys=3)
    images_u8, masks_u8 = synthetic.synthetic_exemplars(11, 3, seed=21)  # chunks of 4, 4, 3 neurons
    kwargs = dict(strategy='rerank', beam=8, group_size=2)
    tokens, scores, steps = engine.describe_host(images_u8.pin_memory(), masks_u8.pin_memory(), **kwargs)
    tokens_p, scores_p, steps_p = engine.describe_host(images_u8, masks_u8, **kwargs)  # pageable host memory
    assert torch.equal(tokens, tokens_p) and torch.equal(scores, scores_p) and torch.equal(steps, steps_p)
    for lo in range(0, 11, 4):
        t, s, st = engine.describe_host(images_u8[lo:lo + 4].contiguous(), masks_u8[lo:lo + 4].contiguous(), **kwargs)
        assert torch.equal(tokens[lo:lo + 4], t) and torch.equal(scores[lo:lo + 4], s) and torch.equal(steps[lo:lo + 4], st)
    # resident inputs: the same pipeline without the copies
    tokens_d, scores_d = engine.describe_device(images_u8.cuda(), masks_u8.cuda(), **kwargs)
    assert torch.equal(tokens_d.cpu(), tokens) and torch.equal(scores_d.cpu(), scores)
    tokens_g, scores_g, steps_g = engine.describe_host(images_u8, masks_u8, strategy='greedy', mi=False)
    feats = engine.encode(images_u8.view(-1, 3, 224, 224), masks_u8.view(-1, 1, 224, 224)).view(11, 3, -1)
    for lo in range(0, 11, 4):
        t, s, _, _ = engine.decode_greedy(feats[lo:lo + 4], 15, False, 0.2)
        assert torch.equal(tokens_g[lo:lo + 4], t.cpu()) and torch.equal(scores_g[lo:lo + 4], s.cpu())
    assert (steps_g == 15).all()
    engine.close()


def test_facade_matches_reference_surface(full_sd):
    """Decoder facade: predict() on a dataset of TopImages-like samples, forward kwargs and error behaviour."""
    from neuron_descriptions_b200 import milan
    from neuron_descriptions_b200.milan import lang
    indexer = lang.Indexer(lang.Vocab(VOCAB), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), max_neurons=16)
    decoder.load_state_dict(full_sd)
    with pytest.raises(RuntimeError):
        decoder(torch.zeros(1, 15, 3904))  # not on a CUDA device: no CPU fallback
    decoder.to('cuda:0')
    images_u8, masks_u8 = synthetic.synthetic_exemplars(3, 15, seed=11)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    dataset = [('layer', i, images_f[i], masks_f[i]) for i in range(3)]
    captions = decoder.predict(dataset, strategy='rerank', temperature=.2, beam_size=50, device='cuda:0',
                               display_progress_as=None)
    with torch.no_grad():
        ref = O.describe(images_f, masks_f, full_sd, VOCAB, strategy='rerank', beam_size=50)
    assert len(captions) == 3 and all(isinstance(c, str) for c in captions)
    assert list(captions) == list(ref)
    out = decoder(images_f, masks_f, strategy='greedy', mi=False)
    assert out.tokens.shape == (3, 15) and out.predictions.shape == (3, 15, V) and out.attentions.shape == (3, 15, 15)
    with pytest.raises(ValueError, match='unknown strategy'):
        decoder(images_f, masks_f, strategy='nope')
    # a longer decode / wider beam than the engine was sized for grows the workspace transparently
    feats = decoder.encode(images_f.cuda(), masks_f.cuda())
    longer = decoder(feats, strategy='beam', mi=False, length=18, beam_size=60)
    ref_long = O.decode(feats.cpu(), full_sd, VOCAB, strategy='beam', mi=False, length=18, beam_size=60)
    assert longer.beam_tokens.shape[-1] == ref_long.beam_tokens.shape[-1]
    torch.testing.assert_close(longer.beam_scores.cpu(), ref_long.beam_scores, atol=LOGP_TOL, rtol=0)
    with pytest.raises(ValueError, match='cannot set `mi=` decoding when reranking'):
        decoder(images_f, masks_f, strategy='rerank', mi=True)
    with pytest.raises(ValueError, match='strategy must have length'):
        decoder(images_f, masks_f, strategy=torch.zeros(3, 4, dtype=torch.long))


def test_score_matches_reference_golden(golden_dir):
    """`Decoder.score` (src/milan/decoders.py:636-711) through the facade vs the reference's own output."""
    from neuron_descriptions_b200 import milan
    from neuron_descriptions_b200.milan import lang
    from oracle.make_golden import SCORE_CAPTIONS, whitespace_tokenize
    g = np.load(os.path.join(golden_dir, 'score.npz'))
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0, with_encoder=False)
    indexer = lang.Indexer(lang.Vocab(VOCAB), tokenize=whitespace_tokenize, start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), max_neurons=4)  # 6 captions -> two engine chunks
    decoder.load_state_dict(sd)
    decoder.to('cuda:0')
    feats = synthetic_features(6, 15, seed=0)
    captions = list(SCORE_CAPTIONS)
    np.testing.assert_allclose(decoder.score(captions, feats, mi=False).cpu().numpy(), g['scores'], atol=LOGP_TOL)
    np.testing.assert_allclose(decoder.score(captions, feats).cpu().numpy(), g['scores_mi'], atol=LOGP_TOL)
    np.testing.assert_allclose(decoder.score(captions, feats[:1], mi=False).cpu().numpy(), g['scores_one'],
                               atol=LOGP_TOL)
    with pytest.raises(ValueError, match='option disallowed'):
        decoder.score(captions, feats, length=3)
    with pytest.raises(ValueError, match='must have batch size 1 or'):
        decoder.score(captions, feats[:2])


def test_edge_cases_vs_oracle():
    """Ragged / extreme sizes: one neuron, beam 1 and the maximum beam, length 1, more keys than the register tile
    of the attention kernel, empty inputs."""
    sd = synthetic.synthetic_state_dict(seed=2, sharpen=12.0, stop_bias=1.0, with_encoder=False)
    from neuron_descriptions_b200.engine import Engine
    engine = Engine(sd, vocab_size=V, device='cuda:0', max_neurons=4, max_beam=64, max_keys=20)
    gen = torch.Generator().manual_seed(5)
    feats = torch.randn(3, 20, synthetic.FEATURE_SIZE, generator=gen).abs() * 0.4  # 20 keys > 16
    ref = O.decode(feats, sd, VOCAB, strategy='greedy', mi=False)
    tok, sc, _, attn = engine.decode_greedy(feats, 15, mi=False, temperature=0.2)
    np.testing.assert_array_equal(tok.cpu().numpy(), ref.tokens.numpy())
    torch.testing.assert_close(sc.cpu(), ref.scores, atol=LOGP_TOL, rtol=0)
    torch.testing.assert_close(attn.cpu(), ref.attentions, atol=1e-4, rtol=0)
    # one neuron, beam 1, length 1 and the largest beam
    one = feats[:1, :15].contiguous()
    for beam, length in ((1, 1), (1, 15), (64, 6)):
        ref = O.decode(one, sd, VOCAB, strategy='rerank', beam_size=beam, length=length)
        bt, bs, steps, tok, sc, _ = engine.decode_beam(one, length, beam, True, 0.2)
        T = int(steps[0])
        assert T == ref.beam_tokens.shape[-1]
        torch.testing.assert_close(bs.cpu(), ref.beam_scores, atol=LOGP_TOL, rtol=0)
        torch.testing.assert_close(sc.cpu(), ref.scores, atol=LOGP_TOL, rtol=0)
        _tokens_match(bt[..., :T].cpu().numpy(), ref.beam_tokens.numpy(), bs.cpu().numpy(), ref.beam_scores.numpy(),
                      f'beam{beam}')
    with pytest.raises(Exception, match='beam size'):
        engine.decode_beam(one, 5, 65, False, 0.2)
    with pytest.raises(Exception, match='exceeds'):
        engine.decode_beam(torch.zeros(5, 15, synthetic.FEATURE_SIZE), 5, 4, False, 0.2)
    engine.close()


def test_encoder_batch_invariance_full_size(full_sd):
    """BASELINE size (64 neurons x 15 exemplars = 960 images per step), size-independent properties of the encoder:
    an image's features do not depend on what else is in the batch (tiles of the implicit GEMM span image boundaries,
    CTA pairs split M tiles between two SMs) - bit for bit -, all-zero masks give exactly zero features, and masked
    pooling is invariant to the image content outside the mask's support."""
    engine = _engine(full_sd, max_neurons=64)
    images_u8, masks_u8 = synthetic.synthetic_exemplars(64, 15, seed=77, zero_mask_fraction=0.05)
    images, masks = images_u8.view(-1, 3, 224, 224), masks_u8.view(-1, 1, 224, 224)
    full = engine.encode(images, masks)
    assert full.shape == (960, synthetic.FEATURE_SIZE) and torch.isfinite(full).all()
    picks = torch.tensor([0, 1, 127, 128, 500, 958, 959])
    alone = engine.encode(images[picks].contiguous(), masks[picks].contiguous())
    assert torch.equal(full[picks], alone), 'features depend on the batch an image is encoded in'
    empty = masks.view(960, -1).sum(dim=1) == 0
    assert empty.any() and (full[empty.to(full.device)] == 0).all()
    # the first retained map (raw conv1, 7x7 receptive field): pixels further than the stem's reach from every mask
    # pixel cannot influence level 0 of the pyramid
    i = int(picks[3])
    keep = torch.nn.functional.max_pool2d(masks[i:i + 1].float(), kernel_size=31, stride=1, padding=15) > 0
    scrambled = images[i:i + 1].clone()
    noise = torch.randint(0, 256, scrambled.shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    scrambled = torch.where(keep.expand_as(scrambled), scrambled, noise)
    again = engine.encode(scrambled, masks[i:i + 1].contiguous())
    assert torch.equal(again[0, :64], full[i, :64])
    engine.close()


def test_empty_inputs(full_engine):
    out = full_engine.encode(torch.zeros(0, 3, 224, 224, dtype=torch.uint8), torch.zeros(0, 1, 224, 224, dtype=torch.uint8))
    assert out.shape == (0, synthetic.FEATURE_SIZE)
    with pytest.raises(ValueError, match='expects'):
        full_engine.encode(torch.zeros(1, 3, 64, 64, dtype=torch.uint8), None)
