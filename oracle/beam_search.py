"""CPU restatement of `allennlp.nn.beam_search.BeamSearch` (allennlp==2.10.0) as MILAN uses it.

TEST INFRASTRUCTURE ONLY (parity oracle). Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module; the product path never does.

PINNING: allennlp is a third-party dependency of the reference (`requirements.txt:7`, pinned 2.10.0) that is
neither vendored under /root/reference nor installed here, and the reference has no test or golden vector for
beam search, so this file cannot be run against the library itself. It restates the
published algorithm of `allennlp/nn/beam_search.py` @ v2.10.0 (`BeamSearch.search/_search`,
`DeterministicSampler`, `SequenceLogProbabilityScorer`) and is pinned by upstream's own published known-answer
tests (`tests/nn/beam_search_test.py::BeamSearchTest`: the 6-state Markov chain, beam 3 ->
[[1,2,3,4,5],[2,3,4,5,5],[3,4,5,5,5]] with log .4/.3/.2, greedy, single step, early stopping, per-node beam
sizes, finished state, bad config, negligible-log-prob and empty-sequence warnings), restated in
`tests/test_beam_search_upstream.py`. The defaults the reference relies on, anchored on its call site
`src/milan/decoders.py:467-484`:

    BeamSearch(end_index=stop_index, max_steps=length, beam_size=beam_size)
        per_node_beam_size = beam_size, sampler = deterministic top-k, min_steps = 0,
        final scorer = summed log-probabilities, no constraints.
    runner.search(start_predictions (B,), start_state: dict, step(tokens, state) -> (log_probs, state))

Further indirect pins on the MILAN decoder itself (tests/test_oracle.py): beam_size=1 reproduces the reference's greedy
decode token-for-token until `<stop>`; every returned beam score equals the forced-decode log-probability sum of
its token sequence; rows come back sorted by score.
"""
import inspect
import warnings
from typing import Callable, Dict, List, Tuple

import torch

StateType = Dict[str, torch.Tensor]


def min_value_of_dtype(dtype: torch.dtype) -> float:
    """allennlp.nn.util.min_value_of_dtype."""
    return torch.finfo(dtype).min


class BeamSearch:
    """Batched deterministic beam search with summed-log-prob scoring."""

    def __init__(self, end_index: int, max_steps: int = 50, beam_size: int = 10, per_node_beam_size: int = None):
        if max_steps <= 0 or beam_size <= 0:
            raise ValueError('max_steps and beam_size must be positive')
        self._end_index = end_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size or beam_size

    @staticmethod
    def _is_multilayer_rnn_decoder(key: str, state_tensor: torch.Tensor) -> bool:
        return state_tensor.dim() == 3 and key in {'decoder_hidden', 'decoder_context'}

    @torch.no_grad()
    def search(self, start_predictions: torch.Tensor, start_state: StateType,
               step: Callable) -> Tuple[torch.Tensor, torch.Tensor]:
        # allennlp accepts 2-argument (tokens, state) step functions and ignores the timestep for them.
        signature = inspect.signature(step)
        if len(signature.parameters) < 3:
            old_step = step

            def new_step(last_predictions, state, time_step):
                del time_step
                return old_step(last_predictions, state)

            return self._search(start_predictions, start_state, new_step)
        return self._search(start_predictions, start_state, step)

    def _search(self, start_predictions, start_state, step):
        batch_size = start_predictions.size()[0]
        predictions: List[torch.Tensor] = []
        backpointers: List[torch.Tensor] = []

        # First step: rows = batch_size.
        start_class_log_probabilities, state = step(start_predictions, start_state, 0)
        num_classes = start_class_log_probabilities.size()[1]
        if self.per_node_beam_size > num_classes:
            raise ValueError(f'Target vocab size ({num_classes:d}) too small relative to per_node_beam_size '
                             f'({self.per_node_beam_size:d}).')

        start_top_log_probabilities, start_predicted_classes = start_class_log_probabilities.topk(self.beam_size)
        if self.beam_size == 1 and (start_predicted_classes == self._end_index).all():
            warnings.warn('Empty sequences predicted. You may want to increase the beam size or ensure your '
                          'step function is working properly.', RuntimeWarning)
            return start_predicted_classes.unsqueeze(-1), start_top_log_probabilities

        last_log_probabilities = start_top_log_probabilities  # (B, beam)
        predictions.append(start_predicted_classes)  # (B, beam)

        # Finished beams may only continue with <stop>, at zero cost.
        log_probs_after_end = start_class_log_probabilities.new_full(
            (batch_size * self.beam_size, num_classes),
            min_value_of_dtype(start_class_log_probabilities.dtype))
        log_probs_after_end[:, self._end_index] = 0.0

        # Expand every state tensor to (B * beam, ...).
        for key, state_tensor in state.items():
            if state_tensor is None:
                continue
            _, *last_dims = state_tensor.size()
            state[key] = (state_tensor.unsqueeze(1).expand(batch_size, self.beam_size, *last_dims).reshape(
                batch_size * self.beam_size, *last_dims))

        for timestep in range(self.max_steps - 1):
            last_predictions = predictions[-1].reshape(batch_size * self.beam_size)
            if (last_predictions == self._end_index).all():
                break

            class_log_probabilities, state = step(last_predictions, state, timestep + 1)

            last_predictions_expanded = last_predictions.unsqueeze(-1).expand(batch_size * self.beam_size,
                                                                              num_classes)
            cleaned_log_probabilities = torch.where(last_predictions_expanded == self._end_index,
                                                    log_probs_after_end, class_log_probabilities)

            top_log_probabilities, predicted_classes = cleaned_log_probabilities.topk(self.per_node_beam_size)

            expanded_last_log_probabilities = (last_log_probabilities.unsqueeze(2).expand(
                batch_size, self.beam_size, self.per_node_beam_size).reshape(batch_size * self.beam_size,
                                                                             self.per_node_beam_size))
            summed_top_log_probabilities = top_log_probabilities + expanded_last_log_probabilities

            reshaped_summed = summed_top_log_probabilities.reshape(batch_size,
                                                                   self.beam_size * self.per_node_beam_size)
            reshaped_predicted_classes = predicted_classes.reshape(batch_size,
                                                                   self.beam_size * self.per_node_beam_size)

            restricted_beam_log_probs, restricted_beam_indices = reshaped_summed.topk(self.beam_size)
            restricted_predicted_classes = reshaped_predicted_classes.gather(1, restricted_beam_indices)
            predictions.append(restricted_predicted_classes)
            last_log_probabilities = restricted_beam_log_probs

            backpointer = torch.divide(restricted_beam_indices, self.per_node_beam_size, rounding_mode='trunc')
            backpointers.append(backpointer)

            # Reorder every state tensor by parent beam.
            for key, state_tensor in state.items():
                if state_tensor is None:
                    continue
                _, *last_dims = state_tensor.size()
                expanded_backpointer = backpointer.view(batch_size, self.beam_size,
                                                        *([1] * len(last_dims))).expand(
                                                            batch_size, self.beam_size, *last_dims)
                state[key] = (state_tensor.reshape(batch_size, self.beam_size, *last_dims).gather(
                    1, expanded_backpointer).reshape(batch_size * self.beam_size, *last_dims))

        if not torch.isfinite(last_log_probabilities).all() or (
                last_log_probabilities == min_value_of_dtype(last_log_probabilities.dtype)).any():
            warnings.warn('Negligible log probabilities encountered (\'-inf\' or equivalent). Some final '
                          'sequences may not make sense.', RuntimeWarning)

        # Backtrack.
        reconstructed_predictions = [predictions[-1].unsqueeze(2)]
        if not backpointers:
            all_predictions = reconstructed_predictions[0]
        else:
            cur_backpointers = backpointers[-1]
            for timestep in range(len(predictions) - 2, 0, -1):
                cur_preds = predictions[timestep].gather(1, cur_backpointers).unsqueeze(2)
                reconstructed_predictions.append(cur_preds)
                cur_backpointers = backpointers[timestep - 1].gather(1, cur_backpointers)
            final_preds = predictions[0].gather(1, cur_backpointers).unsqueeze(2)
            reconstructed_predictions.append(final_preds)
            all_predictions = torch.cat(list(reversed(reconstructed_predictions)), 2)

        # SequenceLogProbabilityScorer: score = summed log-prob; sort each row descending.
        final_scores = last_log_probabilities
        sorted_final_scores, sorted_indices = torch.sort(final_scores, dim=1, descending=True)
        sorted_all_predictions = torch.gather(all_predictions, 1,
                                              sorted_indices.unsqueeze(-1).expand_as(all_predictions))
        return sorted_all_predictions, sorted_final_scores
