"""CPU: the C-ABI shared library loads and exports every symbol include/milan_b200.h declares; the host-side
logic that needs no GPU behaves like the reference."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    header = open(os.path.join(ROOT, 'include', 'milan_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    return sorted(set(re.findall(r'\b(milan_[a-z_0-9]+)\s*\(', header)))


def test_library_exports_every_declared_symbol():
    from neuron_descriptions_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/milan_b200.h but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert lib.milan_version().startswith(b'milan_b200')


def test_config_struct_matches_header():
    from neuron_descriptions_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'milan_b200.h')).read()
    body = header[header.index('typedef struct MilanConfig {'):header.index('} MilanConfig;')]
    fields = re.findall(r'int32_t\s+([a-z_]+);', body)
    assert fields == [name for name, _ in _lib.MilanConfig._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the loud failure without a GPU')
def test_product_path_fails_loudly_without_gpu():
    from neuron_descriptions_b200 import synthetic
    from neuron_descriptions_b200.engine import Engine
    sd = synthetic.synthetic_state_dict(seed=0, with_encoder=False)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Engine(sd, vocab_size=5004, device='cuda:0')


def test_facade_host_logic():
    from neuron_descriptions_b200 import milan, synthetic
    from neuron_descriptions_b200.milan import lang
    vocab = synthetic.synthetic_vocab(50)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    assert (indexer.start_index, indexer.stop_index, indexer.pad_index, indexer.unk_index) == (50, 51, 52, 53)
    assert len(indexer) == 54
    assert indexer.reconstruct([10, 1, 11, 0, 14, 2, 10, 51, 11]) == 'Dog, cat. Top-dog'
    assert indexer.reconstruct([[10], [51, 10]]) == ('Dog', '')
    with pytest.raises(ValueError):
        indexer.reconstruct([])
    with pytest.raises(ValueError):
        indexer.unindex([999])
    with pytest.raises(ValueError, match='encoder not supported'):
        milan.PyramidConvEncoder('vgg16')
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer))
    assert decoder.strategy == 'rerank' and decoder.vocab_size == 54 and decoder.feature_size == 3904
    assert milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101')).strategy == 'beam'
    with pytest.raises(RuntimeError, match='CUDA'):
        decoder(torch.zeros(1, 15, 3904))
    with pytest.raises(KeyError):
        milan.pretrained('not-a-model')
    with pytest.raises(FileNotFoundError):
        milan.pretrained('base', path='/nonexistent/base.pth')


def test_checkpoint_round_trip(tmp_path):
    """Reference payload layout {'properties','children','state_dict'} (src/utils/serialize.py:188-253)."""
    from neuron_descriptions_b200 import milan, synthetic
    from neuron_descriptions_b200.milan import lang
    vocab = synthetic.synthetic_vocab(30)
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), beam_size=7, temperature=.3)
    sd = synthetic.synthetic_state_dict(seed=0, vocab_size=30, with_encoder=False)
    decoder.load_state_dict(sd)
    path = tmp_path / 'milan.pth'
    decoder.save(path)
    payload = torch.load(path, weights_only=False)
    assert set(payload) == {'properties', 'children', 'state_dict'}
    assert payload['children'] == {'encoder': 'PyramidConvEncoder'}
    loaded = milan.Decoder.load(path)
    assert loaded.beam_size == 7 and loaded.temperature == .3 and loaded.lm is not None
    assert loaded.indexer.vocab.tokens == tuple(vocab)
    assert set(loaded.state_dict()) == set(sd)
    via_hub = milan.pretrained('base', path=path)
    assert via_hub.vocab_size == 34


def test_reference_checkpoint_payload_is_ingested(tmp_path):
    """A payload with exactly the structure the UNMODIFIED reference's `Decoder.serialize()` writes
    (`tests/golden/checkpoint_skeleton.json`, made by `oracle/make_golden.py::make_checkpoint_skeleton`; tensors
    materialised from their recorded shapes) loads through `milan.pretrained`, and our own `serialize()` has the
    same structure, so checkpoints move both ways."""
    import json
    from neuron_descriptions_b200 import milan, synthetic
    skeleton = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'checkpoint_skeleton.json')))

    def materialise(node):
        if isinstance(node, list) and len(node) == 3 and node[0] == 'tensor':
            dtype = getattr(torch, node[2].split('.')[-1])
            return torch.zeros(node[1], dtype=dtype) if dtype != torch.float32 else torch.full(node[1], 0.01)
        if isinstance(node, dict):
            return {key: materialise(value) for key, value in node.items()}
        if isinstance(node, list):
            return tuple(materialise(value) for value in node)
        return node

    payload = materialise(skeleton)
    assert set(payload) == {'properties', 'children', 'state_dict'}
    path = tmp_path / 'base.pth'
    torch.save(payload, path)
    decoder = milan.pretrained('base', path=path)
    props = payload['properties']
    assert decoder.length == props['length'] == 15 and decoder.beam_size == props['beam_size'] == 50
    assert decoder.strategy == props['strategy'] == 'rerank' and decoder.temperature == props['temperature']
    assert decoder.indexer.vocab.tokens == tuple(synthetic.synthetic_vocab(40))
    assert decoder.indexer.start_index == 40 and decoder.indexer.stop_index == 41 and decoder.vocab_size == 44
    assert decoder.lm is not None and decoder.lm.hidden_size == 512 and decoder.encoder.config == 'resnet101'
    # every tensor the engine ingests is present under the reference's key names
    state = decoder.state_dict()
    assert len(state) == len(payload['state_dict']) == 658
    for key in ('encoder.mean', 'encoder.encoder.model.layer3.22.conv3.weight', 'lstm.weight_ih', 'output.1.bias',
                'attend.key_to_hidden.weight', 'lm.lstm.weight_hh_l1', 'lm.output.0.weight'):
        assert key in state, key

    # our serializer writes the same structure (the tokenizer slot is None on both sides)
    def structure(node):
        if torch.is_tensor(node):
            return ['tensor', list(node.shape), str(node.dtype)]
        if isinstance(node, dict):
            return {str(key): structure(value) for key, value in node.items()}
        if isinstance(node, (tuple, list)):
            return [structure(value) for value in node]
        return node

    ours = structure(decoder.serialize())
    assert ours['children'] == skeleton['children']
    assert ours['state_dict'] == skeleton['state_dict']
    ref_props, our_props = skeleton['properties'], ours['properties']
    assert set(our_props) == set(ref_props)
    for key in ('embedding_size', 'hidden_size', 'attention_hidden_size', 'dropout', 'length', 'strategy',
                'temperature', 'beam_size'):
        assert our_props[key] == ref_props[key], key
    assert our_props['indexer'] == ref_props['indexer']
    assert our_props['lm']['properties'] == ref_props['lm']['properties']
    assert our_props['encoder']['properties']['config'] == ref_props['encoder']['properties']['config']


def test_checkpoint_with_spacy_tokenizer_child_is_ingested(tmp_path):
    """Real MILAN checkpoints (`milan-base.pth`) carry a serialized spaCy tokenizer inside both indexers:
    `Tokenizer.serialize()` = {'properties': {'nlp': (config_dict, bytes), lemmatize, ...}, 'children': {}}
    (`src/utils/lang.py:14-71`, `src/utils/serialize.py:104-107`; the reference rebuilds the pipeline from the
    tuple at `serialize.py:147-153`). Decoding never tokenises, so the payload must load without spaCy, decode
    ids to text, and survive a save / load round trip untouched."""
    import json
    from neuron_descriptions_b200 import milan
    skeleton = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'checkpoint_skeleton.json')))

    def materialise(node):
        if isinstance(node, list) and len(node) == 3 and node[0] == 'tensor':
            return torch.full(node[1], 0.01, dtype=getattr(torch, node[2].split('.')[-1]))
        if isinstance(node, dict):
            return {key: materialise(value) for key, value in node.items()}
        if isinstance(node, list):
            return tuple(materialise(value) for value in node)
        return node

    payload = materialise(skeleton)
    nlp = ({'nlp': {'lang': 'en', 'pipeline': ['tok2vec', 'tagger', 'lemmatizer']}, 'components': {}},
           b'\x83\xa6config\xc0 not a real spaCy byte string, only its type matters here')
    tokenizer = {'properties': {'nlp': nlp, 'lemmatize': True, 'lowercase': True, 'ignore_stop': True,
                                'ignore_punct': True}, 'children': {}}
    payload['properties']['indexer']['properties']['tokenize'] = tokenizer
    payload['properties']['lm']['properties']['indexer']['properties']['tokenize'] = tokenizer
    path = tmp_path / 'base.pth'
    torch.save(payload, path)
    decoder = milan.pretrained('base', path=path)
    assert decoder.indexer.tokenize is None and decoder.indexer.tokenize_payload == tokenizer
    stop = decoder.indexer.stop_index
    assert decoder.indexer.reconstruct([[5, 0, 6, stop, 7], [stop]]) == ('The. Of', '')  # vocab: . , - ; : the of and
    with pytest.raises(NotImplementedError, match='no tokenizer'):
        decoder.indexer('a caption')  # scoring text needs a tokenizer: pass tokenize= explicitly

    again = tmp_path / 'again.pth'
    decoder.save(again)
    reloaded = torch.load(again, map_location='cpu', weights_only=False)
    for indexer in (reloaded['properties']['indexer'], reloaded['properties']['lm']['properties']['indexer']):
        config, data = indexer['properties']['tokenize']['properties']['nlp']
        assert isinstance(config, dict) and isinstance(data, bytes) and (config, data) == nlp
    assert milan.pretrained('base', path=again).indexer.tokenize_payload == tokenizer
