"""CPU: the host-side `lang.Indexer` mirror (ids <-> text) and the oracle's `score` restatement against goldens
produced by the UNMODIFIED reference (`oracle/make_golden.py::make_score_golden`; reference code:
`src/utils/lang.py:379-515,573-730`, `src/milan/decoders.py:636-711`)."""
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import synthetic
from neuron_descriptions_b200.milan import lang
from oracle import milan_oracle as O
from oracle.make_golden import DEC_NEURONS, K, SCORE_CAPTIONS, synthetic_features, whitespace_tokenize

VOCAB = synthetic.synthetic_vocab(5000)


def _indexer(tokenize=whitespace_tokenize):
    return lang.Indexer(lang.Vocab(VOCAB), tokenize=tokenize, start=True, stop=True, pad=True, unk=True)


def _ragged(rows):
    return np.array([list(row) + [-1] * (16 - len(row)) for row in rows], dtype=np.int64)


def test_indexer_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'score.npz'))
    indexer = _indexer()
    captions = list(SCORE_CAPTIONS)
    np.testing.assert_array_equal(np.array(indexer(captions)), g['indexed_default'])
    np.testing.assert_array_equal(_ragged(indexer(captions, start=False, stop=True, pad=False, unk=True)),
                                  g['indexed_nopad'])
    np.testing.assert_array_equal(np.array(indexer(captions, length=4)), g['indexed_len4'])
    np.testing.assert_array_equal(_ragged(indexer(captions, start=False, stop=False, pad=False, unk=False)),
                                  g['indexed_nounk'])
    np.testing.assert_array_equal(np.array(indexer(captions[2])), g['indexed_single'])
    # the oracle's own restatement agrees too
    np.testing.assert_array_equal(np.array(O.index_tokens(whitespace_tokenize(captions), VOCAB)),
                                  g['indexed_default'])


def test_indexer_round_trip_and_errors():
    indexer = _indexer()
    ids = indexer('dog and zzz cat')
    assert ids[0] == indexer.start_index and ids[-1] == indexer.stop_index and ids[3] == indexer.unk_index
    assert indexer.unindex(ids, specials=False) == ('dog', 'and', 'cat')
    assert indexer.reconstruct(ids) == 'Dog and cat'
    assert indexer.index(()) == ()
    with pytest.raises(NotImplementedError, match='no tokenizer'):
        _indexer(tokenize=None)('dog cat')
    with pytest.raises(ValueError, match='unknown index'):
        indexer.unindex([len(indexer) + 1])


def test_basic_tokenizer():
    tokenize = lang.BasicTokenizer()
    assert tokenize('Dogs, and the Cat\'s toys!') == ('dogs', 'and', 'the', "cat's", 'toys')
    assert tokenize(['A b', '']) == (('a', 'b'), ())


def test_oracle_score_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'score.npz'))
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0, with_encoder=False)
    feats = synthetic_features(DEC_NEURONS, K, seed=0)
    tokenized = whitespace_tokenize(SCORE_CAPTIONS)
    np.testing.assert_allclose(O.score(tokenized, feats, sd, VOCAB, mi=False).numpy(), g['scores'], atol=2e-4)
    np.testing.assert_allclose(O.score(tokenized, feats, sd, VOCAB).numpy(), g['scores_mi'], atol=2e-4)
    np.testing.assert_allclose(O.score(tokenized, feats[:1], sd, VOCAB, mi=False).numpy(), g['scores_one'],
                               atol=2e-4)


def test_spatialize_vit_mlp_matches_reference_semantics():
    """`src/exemplars/transforms.py:55-81`: drop CLS, patches become a square map, units become channels."""
    from neuron_descriptions_b200.exemplars import transforms
    hiddens = torch.arange(2 * 17 * 3, dtype=torch.float32).view(2, 17, 3)
    out = transforms.spatialize_vit_mlp(hiddens)
    assert out.shape == (2, 3, 4, 4)
    for b in range(2):
        for unit in range(3):
            for patch in range(16):
                assert out[b, unit, patch // 4, patch % 4] == hiddens[b, 1 + patch, unit]
    with pytest.raises(AssertionError):
        transforms.spatialize_vit_mlp(torch.zeros(1, 12, 3))


def test_unindex_and_reconstruct_match_reference_golden(golden_dir):
    """`Indexer.unindex` / `reconstruct` vs the unmodified reference on seeded random id sequences
    (`oracle/make_golden.py::make_lang_golden`), for every special-token configuration."""
    import json
    from oracle.make_golden import LANG_FLAGS, LANG_TOKENS, LANG_UNINDEX_KWARGS
    with open(os.path.join(golden_dir, 'lang_reconstruct.json')) as handle:
        cases = json.load(handle)
    assert len(cases) == 40 * len(LANG_FLAGS)
    indexers = [lang.Indexer(lang.Vocab(LANG_TOKENS), tokenize=None, start=a, stop=b, pad=c, unk=d)
                for a, b, c, d in LANG_FLAGS]
    for case in cases:
        indexer, batch = indexers[case['flags']], case['ids']
        for kwargs, want in zip(LANG_UNINDEX_KWARGS, case['unindex']):
            assert [list(seq) for seq in indexer.unindex(batch, **kwargs)] == want
        assert list(indexer.unindex(batch[0])) == case['unindex_single']
        assert list(indexer.reconstruct(batch)) == case['reconstruct']
        assert indexer.reconstruct(batch[0]) == case['reconstruct_single']
        assert list(indexer.reconstruct(indexer.unindex(batch))) == case['reconstruct_tokens']
        assert O.reconstruct(batch[0], LANG_TOKENS) == case['reconstruct_single'] or case['flags'] != 0
    with pytest.raises(ValueError, match='unknown index'):
        indexers[0].unindex([len(LANG_TOKENS) + 4])
    with pytest.raises(ValueError, match='at least one'):
        indexers[0].reconstruct([])
    with pytest.raises(ValueError, match='is empty'):
        indexers[0].reconstruct([[1], []])
