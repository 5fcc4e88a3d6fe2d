"""Facade of `src/milan/lms.py`: the LSTM language model used for the PMI rerank, scored by the CUDA engine."""
from typing import Any, Mapping, Optional

import torch

from neuron_descriptions_b200.milan import lang


class LanguageModel:
    """`src/milan/lms.py:17-131` (inference surface). Bound to an engine by the owning `Decoder`."""

    def __init__(self, indexer: lang.Indexer, embedding_size: int = 128, hidden_size: int = 512, layers: int = 2,
                 dropout: float = .5):
        if layers != 2:
            raise NotImplementedError('milan_b200 implements the 2-layer LM of the shipped checkpoints')
        self.indexer = indexer
        self.embedding_size = embedding_size
        self.hidden_size = hidden_size
        self.layers = layers
        self.dropout = dropout
        self._engine = None

    def bind(self, engine) -> 'LanguageModel':
        self._engine = engine
        return self

    def forward(self, inputs: torch.Tensor, reduce: bool = False, masks: Optional[torch.Tensor] = None):
        """`LanguageModel.forward`, `src/milan/lms.py:58-101`.

        reduce=True with the default mask (everything after the first <stop>, off by one as in `:93-96`) is the rerank
        hot path: one fused scoring call. reduce=False returns the (batch, length, vocab) log-probabilities; a
        caller-supplied mask is applied to the per-token log-probabilities gathered from them, exactly as `:97-100`.
        """
        if self._engine is None:
            raise RuntimeError('language model is not bound to a CUDA engine: call Decoder.to("cuda") first')
        if reduce and masks is None:
            return self._engine.lm_score(inputs)
        inputs = inputs.to(self._engine.device, torch.long)
        rows = max(1, self._engine.max_rows)
        lps = torch.cat([self._engine.lm_logprobs(inputs[lo:lo + rows]) for lo in range(0, len(inputs), rows)])
        if not reduce:
            return lps
        batch_size, length = inputs.shape
        picked = lps[:, :-1].gather(2, inputs[:, 1:].unsqueeze(-1)).squeeze(-1)
        return picked.mul(masks.to(picked.device)).sum(dim=-1)

    __call__ = forward

    def properties(self) -> Mapping[str, Any]:
        return {'indexer': self.indexer, 'embedding_size': self.embedding_size, 'hidden_size': self.hidden_size,
                'layers': self.layers, 'dropout': self.dropout}


def lm(*args, **kwargs):
    raise NotImplementedError('LM training (src/milan/lms.py:283) is out of scope; load a trained checkpoint')
