"""CPU: pins `oracle/beam_search.py` to allennlp's own published known-answer tests.

allennlp==2.10.0 (`requirements.txt:7` of the reference) is neither vendored nor installable here, so the
restatement cannot be run against the library. What CAN be pinned are the known-answer vectors of upstream's
`tests/nn/beam_search_test.py::BeamSearchTest` (v2.10.0): a 6-state Markov chain whose transition matrix, beam
sizes and expected sequences / log-probabilities are part of the published test. The cases below restate that
test (`test_search`, `test_finished_state`, `test_batch_size_of_one`, `test_greedy_search`, `test_single_step`,
`test_early_stopping`, `test_different_per_node_beam_size`, `test_catch_bad_config`,
`test_warn_for_bad_log_probs`, `test_empty_sequences`) against the oracle, i.e. every behaviour of
`BeamSearch.search` the reference's call site (`src/milan/decoders.py:467-484`) relies on. Options the reference
never sets (`min_steps`, samplers, constraints, custom scorers) are outside the restatement.
"""
import numpy as np
import pytest
import torch

from oracle.beam_search import BeamSearch

# Row i = distribution of the token following token i; token 5 is <end>. (upstream `transition_probabilities`)
TRANSITIONS = torch.tensor([
    [0.0, 0.4, 0.3, 0.2, 0.1, 0.0],  # start -> j
    [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],  # 1 -> 2
    [0.0, 0.0, 0.0, 1.0, 0.0, 0.0],  # 2 -> 3
    [0.0, 0.0, 0.0, 0.0, 1.0, 0.0],  # 3 -> 4
    [0.0, 0.0, 0.0, 0.0, 0.0, 1.0],  # 4 -> end
    [0.2, 0.1, 0.2, 0.2, 0.2, 0.3],  # end -> anything (must be ignored: finished beams only emit end)
])
END = TRANSITIONS.shape[0] - 1
EXPECTED_TOP_K = np.array([[1, 2, 3, 4, 5], [2, 3, 4, 5, 5], [3, 4, 5, 5, 5]])
EXPECTED_LOG_PROBS = np.log(np.array([0.4, 0.3, 0.2]))


def take_step_no_timestep(last_predictions, state):
    """2-argument step, like the closure the reference passes (`decoders.py:471-481`)."""
    rows = [torch.log(TRANSITIONS[int(tok)]) for tok in last_predictions]
    return torch.stack(rows), state


def take_step_with_timestep(last_predictions, state, timestep):
    return take_step_no_timestep(last_predictions, state)


def check(beam_search=None, batch_size=5, expected_top_k=EXPECTED_TOP_K, expected_log_probs=EXPECTED_LOG_PROBS,
          state=None, take_step=take_step_with_timestep):
    beam_search = beam_search or BeamSearch(END, max_steps=10, beam_size=3)
    state = {} if state is None else state
    beam_size = beam_search.beam_size
    top_k, log_probs = beam_search.search(torch.tensor([0] * batch_size), state, take_step)
    assert list(top_k.size())[:-1] == [batch_size, beam_size]
    assert list(log_probs.size()) == [batch_size, beam_size]
    for b in range(batch_size):  # upstream checks item 0; every batch item is identical here
        np.testing.assert_array_equal(top_k[b].numpy(), expected_top_k)
        np.testing.assert_allclose(log_probs[b].numpy(), expected_log_probs, rtol=1e-6)


@pytest.mark.parametrize('step_fn', [take_step_with_timestep, take_step_no_timestep])
def test_search(step_fn):
    check(take_step=step_fn)


def test_finished_state():
    """State tensors are expanded to (batch * beam, ...) and follow the backpointers; the dict is updated in place."""
    foo = [[1, 0, 1], [2, 0, 1], [0, 0, 1], [1, 1, 1], [0, 0, 0]]
    state = {'foo': torch.tensor(foo)}
    check(state=state)
    expected = np.repeat(np.array(foo), 3, axis=0)
    np.testing.assert_array_equal(state['foo'].numpy(), expected)


def test_batch_size_of_one():
    check(batch_size=1)


def test_greedy_search():
    check(beam_search=BeamSearch(END, beam_size=1), expected_top_k=np.array([[1, 2, 3, 4, 5]]),
          expected_log_probs=np.log(np.array([0.4])))


def test_single_step():
    check(beam_search=BeamSearch(END, max_steps=1, beam_size=3), expected_top_k=np.array([[1], [2], [3]]),
          expected_log_probs=np.log(np.array([0.4, 0.3, 0.2])))


def test_early_stopping():
    """`max_steps` reached before any beam ends: sequences are cut, scores are the sums so far."""
    check(beam_search=BeamSearch(END, beam_size=3, max_steps=3),
          expected_top_k=np.array([[1, 2, 3], [2, 3, 4], [3, 4, 5]]),
          expected_log_probs=np.log(np.array([0.4, 0.3, 0.2])))


@pytest.mark.parametrize('per_node', [1, 2])
def test_different_per_node_beam_size(per_node):
    check(beam_search=BeamSearch(END, beam_size=3, per_node_beam_size=per_node))


def test_catch_bad_config():
    """per_node_beam_size (= beam_size) larger than the number of classes is a configuration error."""
    with pytest.raises(ValueError, match='too small relative to per_node_beam_size'):
        BeamSearch(END, beam_size=20).search(torch.tensor([0] * 5), {}, take_step_with_timestep)


def test_warn_for_bad_log_probs():
    """From token 4 the only continuation is <end>; a beam of 3 must pick 2 zero-probability beams -> warning."""
    initial = torch.LongTensor([END - 1, END - 1])
    with pytest.warns(RuntimeWarning, match='Negligible log probabilities'):
        BeamSearch(END, max_steps=10, beam_size=3).search(initial, {}, take_step_with_timestep)


def test_empty_sequences():
    initial = torch.LongTensor([END - 1, END - 1])
    with pytest.warns(RuntimeWarning, match='Empty sequences predicted'):
        predictions, log_probs = BeamSearch(END, beam_size=1).search(initial, {}, take_step_with_timestep)
    assert list(predictions.size()) == [2, 1, 1]
    assert list(log_probs.size()) == [2, 1]
    assert (predictions == END).all()
    assert (log_probs == 0).all()


def test_finished_beams_only_emit_end():
    """Once a beam has produced <end> its row of the step output is replaced: the `end -> anything` row of the
    transition matrix (which would prefer token 5 at 0.3 but also offers 0..4) never contributes."""
    top_k, log_probs = BeamSearch(END, max_steps=10, beam_size=3).search(torch.tensor([0]), {}, take_step_no_timestep)
    for seq in top_k[0].tolist():
        first_end = seq.index(END)
        assert all(tok == END for tok in seq[first_end:])
    np.testing.assert_allclose(log_probs[0].numpy(), EXPECTED_LOG_PROBS, rtol=1e-6)  # stop-padding costs 0
