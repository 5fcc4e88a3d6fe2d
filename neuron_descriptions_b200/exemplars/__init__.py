"""Stage-1 exemplar computation on B200 (mirror of `src/exemplars/__init__.py:15`): `discriminative`, `generative`,
`compute`."""
from neuron_descriptions_b200.exemplars.compute import ActivationStats, compute, discriminative, generative
from neuron_descriptions_b200.exemplars import transforms  # noqa: E402,F401
