"""Import the UNMODIFIED reference (`/root/reference/src/milan`) in this container.

TEST INFRASTRUCTURE ONLY. Used by `oracle/make_golden.py` (and `tests/test_oracle_vs_reference.py` when
`/root/reference` exists) to pin the oracle restatement against the reference's own code. `/root/reference`
does not exist on the GPU box, so nothing imported at run time by `-m gpu` tests, `smoke()` or `bench.py`
may depend on this module.

The reference imports seven third-party packages that are not installed here;
none of them is executed on the describe-neurons path except `allennlp.nn.beam_search.BeamSearch`, for which
we plug in the restatement from `oracle/beam_search.py` (allennlp==2.10 is not vendored in the reference).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('MILAN_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src', 'milan'))


class _AttrDict(dict):
    """Stand-in for easydict.EasyDict (dict with attribute access; accepts d=...)."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        for key, value in dict(d or {}, **kwargs).items():
            self[key] = value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as error:
            raise AttributeError(key) from error

    def __setattr__(self, key, value):
        self[key] = value


def _stub(name, **attrs):
    module = types.ModuleType(name)
    for key, value in attrs.items():
        setattr(module, key, value)
    sys.modules[name] = module
    return module


def install_stubs():
    """Pre-seed sys.modules with placeholders for the reference's absent third-party imports."""
    from oracle import beam_search as restated

    class _Placeholder:  # never instantiated on the hot path
        def __init__(self, *args, **kwargs):
            raise RuntimeError('placeholder for an uninstalled third-party class')

    if 'bert_score' not in sys.modules:
        _stub('bert_score', BERTScorer=_Placeholder)
    if 'sacrebleu' not in sys.modules:
        _stub('sacrebleu', BLEUScore=_Placeholder, corpus_bleu=None)
    if 'rouge' not in sys.modules:
        _stub('rouge', Rouge=_Placeholder)
    if 'clip' not in sys.modules:
        _stub('clip')
    if 'spacy' not in sys.modules:
        spacy = _stub('spacy', Language=_Placeholder, load=None, util=types.SimpleNamespace())
        lang = _stub('spacy.lang')
        en = _stub('spacy.lang.en', English=_Placeholder)
        spacy.lang = lang
        lang.en = en
    # the stage-1 path (src/exemplars -> src/deps/netdissect) additionally imports these; nothing of them is executed
    if 'statsmodels' not in sys.modules:
        _stub('statsmodels')
        _stub('statsmodels.stats')
        _stub('statsmodels.stats.correlation_tools', cov_nearest=None, corr_nearest=None)
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _stub('matplotlib', cm=types.SimpleNamespace(hot=None))
    if 'easydict' not in sys.modules:
        _stub('easydict', EasyDict=_AttrDict)
    if 'allennlp' not in sys.modules:
        allennlp = _stub('allennlp')
        nn = _stub('allennlp.nn')
        bs = _stub('allennlp.nn.beam_search', BeamSearch=restated.BeamSearch)
        allennlp.nn = nn
        nn.beam_search = bs


def import_reference():
    """Return the reference's (milan, lang) modules, imported from REFERENCE_ROOT."""
    if not reference_available():
        raise RuntimeError(f'reference not found at {REFERENCE_ROOT}')
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # The reference package is called `src`; make sure we do not pick up this repo's compatibility shim.
    for name in [m for m in sys.modules if m == 'src' or m.startswith('src.')]:
        module = sys.modules[name]
        path = getattr(module, '__file__', '') or ''
        if not path.startswith(REFERENCE_ROOT):
            del sys.modules[name]
    import importlib
    milan = importlib.import_module('src.milan')
    lang = importlib.import_module('src.utils.lang')
    return milan, lang


def import_reference_exemplars():
    """The reference's stage-1 module (`src/exemplars/compute.py`), imported with the same stubs."""
    import_reference()
    import importlib
    return importlib.import_module('src.exemplars.compute')
