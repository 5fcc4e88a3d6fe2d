"""Reranking of MILAN's beam by an image-text similarity model (facade of `src/milan/rerankers.py`).

The reference's reranker is CLIP ViT-B/32 with the CLS-token attention of its visual encoder edited by the activation
mask (`rerankers.py:36-250`). Neither the `clip` package nor its weights exist offline, so the MODEL is not rebuilt
here; what is, is everything around it — the reranking procedure (`CLIPWithMasksReranker.forward`, `:261-330`) and
`DecoderWithCLIP` (`decoders.py:1115-1211`) — written against the one call the procedure makes of the model:

    similarity(images (k, 3, H, W), texts (n strings), masks=None | (k, 1, H, W)) -> (k, n) scores

which is `CLIPWithMasks.forward`'s contract (`:141-229`). Hand `SimilarityReranker` any such callable (the
reference's own `CLIPWithMasks` instance works unchanged). Pinned by `tests/golden/reranker.json`, produced by the
UNMODIFIED reference class around a deterministic stand-in similarity (`oracle/make_golden.py::make_reranker_golden`).
"""
from typing import Any, Callable, NamedTuple, Optional, Sequence

import torch

StrSequence = Sequence[str]


class RerankerOutput(NamedTuple):
    """`rerankers.py:253-258`."""
    texts: Sequence[StrSequence]
    orders: Sequence[Sequence[int]]
    scores: Sequence[Sequence[float]]


class SimilarityReranker:
    """`CLIPWithMasksReranker` (`rerankers.py:261-330`) around any `similarity(images, texts, masks=None)`."""

    def __init__(self, similarity: Callable[..., torch.Tensor], lam: float = .5):
        self.similarity = similarity
        self.lam = lam

    def forward(self, images: torch.Tensor, masks: torch.Tensor, texts: Sequence[StrSequence],
                lam: Optional[float] = None) -> RerankerOutput:
        """For every neuron: score each candidate text against each exemplar with and without its mask, sum over the
        exemplars, mix `(1 - lam) * masked + lam * unmasked`, sort descending."""
        if len(images) != len(masks):
            raise ValueError('images and masks batch sizes do not align: '
                             f'{len(images)} vs. {len(masks)}')
        if len(images) != len(texts):
            raise ValueError('images and texts batch sizes do not align: '
                             f'{len(images)} vs. {len(texts)}')
        lam = self.lam if lam is None else lam
        ranked, orders, scores = [], [], []
        for neuron_images, neuron_masks, candidates in zip(images, masks, texts):
            with_masks = self.similarity(neuron_images, candidates, masks=neuron_masks).sum(dim=0)
            without = self.similarity(neuron_images, candidates).sum(dim=0)
            values, order = ((1. - lam) * with_masks + lam * without).sort(descending=True)
            order = order.tolist()
            ranked.append(tuple(candidates[i] for i in order))
            orders.append(tuple(order))
            scores.append(tuple(values.tolist()))
        return RerankerOutput(tuple(ranked), tuple(orders), tuple(scores))

    __call__ = forward


def reranker(lam: float = 1., similarity: Optional[Callable[..., torch.Tensor]] = None, **kwargs: Any) -> SimilarityReranker:
    """`rerankers.reranker` (`:333-345`). The reference builds CLIP here (`clip.load(name)`); offline the similarity
    model has to be supplied."""
    if similarity is None:
        raise NotImplementedError(
            'the CLIP reranker needs the `clip` package and its ViT-B/32 weights, neither of which is available '
            'offline: pass similarity=<callable (images, texts, masks=None) -> (k, n) scores>, e.g. the reference\'s '
            f'CLIPWithMasks instance (ignored arguments: {sorted(kwargs)})')
    return SimilarityReranker(similarity, lam=lam)
