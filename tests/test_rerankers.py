"""CPU: `milan.rerankers.SimilarityReranker` vs the golden produced by the UNMODIFIED reference's
`CLIPWithMasksReranker.forward` (`src/milan/rerankers.py:261-330`) around the same stand-in similarity model."""
import json
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200.milan import rerankers
from oracle.make_golden import reranker_inputs, reranker_similarity


@pytest.mark.parametrize('case', ['default', 'lam0.2', 'unmasked_only'])
def test_similarity_reranker_matches_reference_golden(golden_dir, case):
    with open(os.path.join(golden_dir, 'reranker.json')) as handle:
        g = json.load(handle)[case]
    images, masks, texts = reranker_inputs()
    reranker = rerankers.SimilarityReranker(reranker_similarity, lam=g['default_lam'])
    got = reranker(images, masks, texts, lam=g['lam'])
    assert [list(o) for o in got.orders] == g['orders']
    assert [list(t) for t in got.texts] == g['texts']
    for mine, ref in zip(got.scores, g['scores']):
        np.testing.assert_allclose(mine, ref, rtol=1e-6, atol=1e-6)


def test_reranker_argument_checks():
    images, masks, texts = reranker_inputs()
    reranker = rerankers.SimilarityReranker(reranker_similarity)
    with pytest.raises(ValueError, match='images and masks batch sizes do not align'):
        reranker(images, masks[:2], texts)
    with pytest.raises(ValueError, match='images and texts batch sizes do not align'):
        reranker(images, masks, texts[:1])
    with pytest.raises(NotImplementedError, match='clip'):
        rerankers.reranker()  # the reference would download CLIP ViT-B/32 here
    assert isinstance(rerankers.reranker(lam=.3, similarity=reranker_similarity), rerankers.SimilarityReranker)
    unmasked_only = rerankers.reranker(similarity=reranker_similarity)  # default lam = 1 (`rerankers.py:333`)
    ref = rerankers.SimilarityReranker(reranker_similarity, lam=1.)
    assert unmasked_only(images, masks, texts).orders == ref(images, masks, texts).orders
    assert torch.is_tensor(reranker_similarity(images[0], texts[0]))
