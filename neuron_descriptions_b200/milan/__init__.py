"""MILAN on B200: same exports as `src/milan/__init__.py:13-17` (inference surface)."""
from neuron_descriptions_b200.milan.decoders import (Decoder, DecoderOutput, DecoderState, DecoderStep,
                                                     DecoderWithCLIP, decoder)
from neuron_descriptions_b200.milan import rerankers  # noqa: F401
from neuron_descriptions_b200.milan.encoders import (Encoder, PyramidConvEncoder, SpatialConvEncoder, encoder)
from neuron_descriptions_b200.milan.lms import LanguageModel, lm
from neuron_descriptions_b200.milan.loaders import pretrained
