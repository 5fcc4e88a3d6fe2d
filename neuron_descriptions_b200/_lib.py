"""ctypes binding of libmilan_b200.so (C ABI in include/milan_b200.h).

There is deliberately NO fallback: if the shared library is missing or no B200 is present, the product path
raises. Build with `python -c "import __graft_entry__ as g; g.build()"` or `make -C neuron_descriptions_b200/csrc`.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmilan_b200.so')

PRECISION_SPLIT, PRECISION_FAST = 0, 1
DTYPE_U8, DTYPE_F32 = 0, 1
STRATEGY_GREEDY, STRATEGY_BEAM, STRATEGY_RERANK = 0, 1, 2
ENCODER_ARCHS = {'resnet101': 0, 'resnet50': 1, 'resnet18': 2, 'resnet34': 3, 'alexnet': 4}
ENCODER_KINDS = {'pyramid': 0, 'spatial': 1}


class MilanConfig(ctypes.Structure):
    """Mirror of `struct MilanConfig` (include/milan_b200.h)."""
    _fields_ = [(name, c_int32) for name in (
        'vocab_size', 'embedding_size', 'hidden_size', 'attention_size', 'feature_size', 'start_index',
        'stop_index', 'has_encoder', 'has_lm', 'lm_embedding_size', 'lm_hidden_size', 'precision', 'max_images',
        'max_neurons', 'max_beam', 'max_keys', 'max_length', 'encoder_arch', 'encoder_kind')]


# name -> (restype, argtypes); must list every function declared in include/milan_b200.h
SIGNATURES = {
    'milan_version': (c_char_p, []),
    'milan_last_error': (c_char_p, []),
    'milan_engine_create': (c_int32, [POINTER(MilanConfig), c_int32, POINTER(c_void_p)]),
    'milan_engine_destroy': (None, [c_void_p]),
    'milan_engine_set_tensor': (c_int32, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32]),
    'milan_engine_finalize': (c_int32, [c_void_p]),
    'milan_encode': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    'milan_init_state': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    'milan_step': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                             c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    'milan_decode_greedy': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'milan_decode_beam': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'milan_lm_score': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    'milan_lm_logprobs': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    'milan_describe_host': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'milan_describe_device': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    'milan_tally_topk': (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    'milan_tally_samples': (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64, c_int64, c_void_p]),
    'milan_tally_hist': (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'milan_quantile_exact': (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    'milan_quantile_hist': (c_int32, [c_void_p, c_int32, c_int64, c_float, c_void_p, c_void_p]),
    'milan_activation_masks': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'milan_launch_count': (c_int64, []),
    'milan_set_profiling': (c_int32, [c_void_p, c_int32]),
    'milan_get_profile': (c_int32, [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float),
                                    POINTER(c_int64)]),
}

_lib = None


def load():
    """Load the shared library (once) and declare its signatures. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: the milan_b200 engine is CUDA-only and has no fallback. '
            'Build it with `make -C neuron_descriptions_b200/csrc` (or __graft_entry__.build()).')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class MilanError(RuntimeError):
    """A non-zero status from the C ABI."""


def check(status: int):
    if status != 0:
        message = load().milan_last_error().decode('utf-8', 'replace')
        # Argument errors mirror the reference's ValueErrors (src/milan/decoders.py:395-409, :605-608).
        if any(key in message for key in ('cannot use MI', 'cannot set `mi=`', 'state must have', 'state has h_lm',
                                          'too small relative')):
            raise ValueError(message)
        raise MilanError(message)
