"""CPU (torch fp32) restatement of the reference's describe-neurons path.

TEST INFRASTRUCTURE ONLY (parity oracle). Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module. The product path
(`neuron_descriptions_b200/`) never imports anything under `oracle/` and fails loudly without its CUDA library.

Every function restates one reference function on plain tensors + a flat reference-format state dict (the
`state_dict` of a reference `Decoder`, key names in SURVEY.md section 5), citing the reference file:line it
follows. Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (imported from /root/reference with
stubs for absent third-party packages) on seeded inputs and commits its outputs under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against those vectors. The beam search is the one exception
(third-party allennlp, see `oracle/beam_search.py`: parity unpinned, indirect pins only).
"""
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from oracle.beam_search import BeamSearch

State = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # torchvision BatchNorm2d default


# ----------------------------------------------------------------------------------------------- encoder
def _bn(x: torch.Tensor, sd: State, prefix: str) -> torch.Tensor:
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], training=False, eps=BN_EPS)


def _bottleneck(x: torch.Tensor, sd: State, prefix: str, stride: int) -> torch.Tensor:
    """torchvision Bottleneck v1.5 (stride on the 3x3), eval mode."""
    identity = x
    out = F.relu(_bn(F.conv2d(x, sd[prefix + '.conv1.weight']), sd, prefix + '.bn1'))
    out = F.relu(_bn(F.conv2d(out, sd[prefix + '.conv2.weight'], stride=stride, padding=1), sd, prefix + '.bn2'))
    out = _bn(F.conv2d(out, sd[prefix + '.conv3.weight']), sd, prefix + '.bn3')
    if prefix + '.downsample.0.weight' in sd:
        identity = _bn(F.conv2d(x, sd[prefix + '.downsample.0.weight'], stride=stride), sd,
                       prefix + '.downsample.1')
    return F.relu(out + identity)


def _basic_block(x: torch.Tensor, sd: State, prefix: str, stride: int) -> torch.Tensor:
    """torchvision BasicBlock (resnet18/34), eval mode."""
    identity = x
    out = F.relu(_bn(F.conv2d(x, sd[prefix + '.conv1.weight'], stride=stride, padding=1), sd, prefix + '.bn1'))
    out = _bn(F.conv2d(out, sd[prefix + '.conv2.weight'], padding=1), sd, prefix + '.bn2')
    if prefix + '.downsample.0.weight' in sd:
        identity = _bn(F.conv2d(x, sd[prefix + '.downsample.0.weight'], stride=stride), sd,
                       prefix + '.downsample.1')
    return F.relu(out + identity)


RESNET_ARCHS = {'resnet101': (True, (3, 4, 23, 3)), 'resnet50': (True, (3, 4, 6, 3)),
                'resnet18': (False, (2, 2, 2, 2)), 'resnet34': (False, (3, 4, 6, 3))}


def resnet_retained(images: torch.Tensor, sd: State, prefix: str = 'encoder.encoder.model.',
                    arch: str = 'resnet101') -> List[torch.Tensor]:
    """Outputs of modules ('conv1','layer1','layer2','layer3','layer4') as nethook retains them.

    Follows `src/milan/encoders.py:273-276,298-299` + `src/deps/netdissect/nethook.py:226-235`: 'conv1' is the
    RAW 7x7 conv output (before bn1/ReLU); layerN are post-residual-ReLU stage outputs. `arch` picks the
    torchvision graph of `PyramidConvEncoder.configs()` (`src/milan/encoders.py:326-351`).
    """
    bottleneck, stage_blocks = RESNET_ARCHS[arch]
    block = _bottleneck if bottleneck else _basic_block
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    retained = []
    x = F.conv2d(images, sub['conv1.weight'], stride=2, padding=3)
    retained.append(x)
    x = F.relu(_bn(x, sub, 'bn1'))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, blocks in enumerate(stage_blocks, start=1):
        for bi in range(blocks):
            stride = 2 if (bi == 0 and li > 1) else 1
            x = block(x, sub, f'layer{li}.{bi}', stride)
        retained.append(x)
    return retained


def resnet101_retained(images: torch.Tensor, sd: State, prefix: str = 'encoder.encoder.model.') -> List[torch.Tensor]:
    return resnet_retained(images, sd, prefix, 'resnet101')


def alexnet_retained(images: torch.Tensor, sd: State, prefix: str = 'encoder.encoder.model.') -> List[torch.Tensor]:
    """Outputs of torchvision alexnet `features.{0,3,6,8,10}` as nethook retains them (`encoders.py:328-334`).

    nethook stores `output.detach()` (`nethook.py:226-235`), which shares storage with the conv output, and every
    following `nn.ReLU(inplace=True)` of torchvision's alexnet then rectifies that storage: the retained maps are
    the POST-ReLU activations (verified against the reference in `oracle/make_golden.py`).
    """
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    retained = []
    x = F.relu(F.conv2d(images, sub['features.0.weight'], sub['features.0.bias'], stride=4, padding=2))
    retained.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2)
    x = F.relu(F.conv2d(x, sub['features.3.weight'], sub['features.3.bias'], padding=2))
    retained.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2)
    for name in ('features.6', 'features.8', 'features.10'):
        x = F.relu(F.conv2d(x, sub[name + '.weight'], sub[name + '.bias'], padding=1))
        retained.append(x)
    return retained


def masked_pool(features: Sequence[torch.Tensor], masks: torch.Tensor) -> torch.Tensor:
    """`src/milan/encoders.py:301-320`: bilinear mask downsample, per-image sum-normalise, weighted pool."""
    masked = []
    for fs in features:
        ms = F.interpolate(masks, size=fs.shape[-2:], mode='bilinear', align_corners=False)
        zeros = torch.zeros_like(ms)
        valid = ~ms.isclose(zeros).all(dim=-1).all(dim=-1).view(-1)
        indices = valid.nonzero().squeeze()
        ms[indices] /= ms[indices].sum(dim=(-1, -2), keepdim=True)
        masked.append(fs.mul(ms).sum(dim=(-1, -2)))
    return torch.cat(masked, dim=-1)


def pyramid_encode(images: torch.Tensor, masks: Optional[torch.Tensor], sd: State,
                   arch: str = 'resnet101') -> torch.Tensor:
    """`PyramidConvEncoder.forward`, `src/milan/encoders.py:286-320`. images (N,3,H,W) in [0,1]."""
    if masks is None:
        masks = images.new_ones((len(images), 1, *images.shape[2:]))
    images = (images - sd['encoder.mean']) / sd['encoder.std']
    retained = alexnet_retained(images, sd) if arch == 'alexnet' else resnet_retained(images, sd, arch=arch)
    return masked_pool(retained, masks.clone())


def spatial_encode(images: torch.Tensor, masks: Optional[torch.Tensor], sd: State,
                   arch: str = 'resnet18') -> torch.Tensor:
    """`SpatialConvEncoder.forward`, `src/milan/encoders.py:193-214`: (N,3,H,W) -> (N, 49, 512)."""
    if masks is None:
        masks = images.new_ones((len(images), 1, *images.shape[2:]))
    images = (images - sd['encoder.mean']) / sd['encoder.std']
    features = resnet_retained(images * masks, sd, arch=arch)[-1]
    features = features.permute(0, 2, 3, 1)
    return features.reshape(len(images), -1, features.shape[-1])


def encode(images: torch.Tensor, masks: Optional[torch.Tensor], sd: State, arch: str = 'resnet101',
           kind: str = 'pyramid') -> torch.Tensor:
    """`Decoder.encode`, `src/milan/decoders.py:525-546`: (B,k,3,H,W) -> (B,n_keys,F)."""
    batch_size = len(images)
    images = images.reshape(-1, *images.shape[-3:])
    if masks is not None:
        masks = masks.reshape(-1, *masks.shape[-3:])
    if kind == 'spatial':
        features = spatial_encode(images, masks, sd, arch)
    else:
        features = pyramid_encode(images, masks, sd, arch)
    return features.reshape(batch_size, -1, features.shape[-1])


# ----------------------------------------------------------------------------------------------- decoder
class DecoderState(NamedTuple):
    h: torch.Tensor
    c: torch.Tensor
    h_lm: Optional[torch.Tensor]
    c_lm: Optional[torch.Tensor]


def init_state(features: torch.Tensor, sd: State, lm: bool = False) -> DecoderState:
    """`Decoder.init_state`, `src/milan/decoders.py:548-574`."""
    pooled = features.mean(dim=1)
    h = torch.tanh(F.linear(pooled, sd['init_h.0.weight'], sd['init_h.0.bias']))
    c = torch.tanh(F.linear(pooled, sd['init_c.0.weight'], sd['init_c.0.bias']))
    h_lm = c_lm = None
    if lm:
        hidden = sd['lm.lstm.weight_hh_l0'].shape[1]
        h_lm = h.new_zeros(2, len(features), hidden)
        c_lm = h.new_zeros(2, len(features), hidden)
    return DecoderState(h, c, h_lm, c_lm)


def attention(query: torch.Tensor, keys: torch.Tensor, sd: State) -> torch.Tensor:
    """`Attention.forward`, `src/milan/decoders.py:57-73` (softmax over keys, dim=1)."""
    q_hidden = F.linear(query, sd['attend.query_to_hidden.weight'], sd['attend.query_to_hidden.bias']).unsqueeze(1)
    k_hidden = F.linear(keys, sd['attend.key_to_hidden.weight'], sd['attend.key_to_hidden.bias'])
    hidden = torch.tanh(q_hidden + k_hidden)
    scores = F.linear(hidden, sd['attend.output.0.weight'], sd['attend.output.0.bias'])
    return torch.softmax(scores, dim=1).view(*keys.shape[:2])


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell semantics: gate order i, f, g, o."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=-1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def lm_step(tokens: torch.Tensor, h_lm: torch.Tensor, c_lm: torch.Tensor, sd: State):
    """One token through the LM's embedding + 2-layer LSTM + output (`src/milan/decoders.py:624-630`)."""
    x = F.embedding(tokens, sd['lm.embedding.weight'])
    hs, cs = [], []
    for layer in range(2):
        h, c = lstm_cell(x, h_lm[layer], c_lm[layer], sd[f'lm.lstm.weight_ih_l{layer}'],
                         sd[f'lm.lstm.weight_hh_l{layer}'], sd[f'lm.lstm.bias_ih_l{layer}'],
                         sd[f'lm.lstm.bias_hh_l{layer}'])
        hs.append(h)
        cs.append(c)
        x = h  # inter-layer dropout inactive in eval (src/milan/lms.py:50-54)
    log_p = torch.log_softmax(F.linear(x, sd['lm.output.0.weight'], sd['lm.output.0.bias']), dim=-1)
    return log_p, torch.stack(hs), torch.stack(cs)


class DecoderStep(NamedTuple):
    predictions: torch.Tensor
    attentions: torch.Tensor
    state: DecoderState


def step(features: torch.Tensor, tokens: torch.Tensor, state: DecoderState, sd: State,
         temperature: float = 0.2) -> DecoderStep:
    """`Decoder.step`, `src/milan/decoders.py:576-634` (eval mode: dropout inactive)."""
    h, c, h_lm, c_lm = state
    attentions = attention(h, features, sd)
    attenuated = attentions.unsqueeze(-1).mul(features).sum(dim=1)
    gate = torch.sigmoid(F.linear(h, sd['feature_gate.0.weight'], sd['feature_gate.0.bias']))
    gated = attenuated * gate
    embeddings = F.embedding(tokens, sd['embedding.weight'])
    inputs = torch.cat((embeddings, gated), dim=-1)
    h, c = lstm_cell(inputs, h, c, sd['lstm.weight_ih'], sd['lstm.weight_hh'], sd['lstm.bias_ih'], sd['lstm.bias_hh'])
    predictions = torch.log_softmax(F.linear(h, sd['output.1.weight'], sd['output.1.bias']), dim=-1)
    if h_lm is not None and c_lm is not None:
        log_p_lm, h_lm, c_lm = lm_step(tokens, h_lm, c_lm, sd)
        predictions = predictions - temperature * log_p_lm
    return DecoderStep(predictions, attentions, DecoderState(h, c, h_lm, c_lm))


def lm_forward(inputs: torch.Tensor, sd: State, stop_index: int, reduce: bool = True) -> torch.Tensor:
    """`LanguageModel.forward`, `src/milan/lms.py:58-101`, including the stop-mask off-by-one (:93-96)."""
    batch_size, length = inputs.shape
    h = inputs.new_zeros(2, batch_size, sd['lm.lstm.weight_hh_l0'].shape[1], dtype=torch.float32)
    c = torch.zeros_like(h)
    lps = []
    for t in range(length):
        log_p, h, c = lm_step(inputs[:, t], h, c, sd)
        lps.append(log_p)
    lps = torch.stack(lps, dim=1)  # (B, length, V)
    if not reduce:
        return lps
    idx_batch = torch.arange(batch_size).repeat_interleave(length - 1)
    idx_time = torch.arange(length - 1).repeat(batch_size)
    idx_tokens = inputs[:, 1:].reshape(-1)
    masks = inputs.new_ones((batch_size, length - 1))
    for i, j in inputs.eq(stop_index).nonzero():
        masks[i, j + 1:] = 0
    return lps[:, :-1][idx_batch, idx_time, idx_tokens].view(batch_size, length - 1).mul(masks).sum(dim=-1)


def reconstruct(tokens: Sequence[int], vocab: Sequence[str]) -> str:
    """`Indexer.reconstruct` + `unindex`, `src/utils/lang.py:573-612,678-730`, for one id sequence."""
    n = len(vocab)
    specials = {n: '<start>', n + 1: '<stop>', n + 2: '<pad>', n + 3: '<unk>'}
    words = []
    for index in tokens:
        if index < n:
            words.append(vocab[index])
        elif index in specials:
            words.append(specials[index])
        else:
            raise ValueError(f'unknown index: {index}')
    if '<stop>' in words:
        words = words[:words.index('<stop>')]
    text = ' '.join(w for w in words if w not in specials.values())
    for token in ('.', ',', ';', ':'):
        text = text.replace(' ' + token, token)
    for token in ('-',):
        text = text.replace(' %s' % token, token)
        text = text.replace('%s ' % token, token)
    return '. '.join(sentence.strip().capitalize() for sentence in text.split('.')).strip()


class DecoderOutput(NamedTuple):
    captions: Tuple[str, ...]
    scores: torch.Tensor
    tokens: torch.Tensor
    predictions: Optional[torch.Tensor]
    attentions: Optional[torch.Tensor]
    beam_scores: Optional[torch.Tensor]
    beam_tokens: Optional[torch.Tensor]


@torch.no_grad()
def decode(features: torch.Tensor, sd: State, vocab: Sequence[str], strategy: str = 'rerank', length: int = 15,
           temperature: float = 0.2, beam_size: int = 50, mi: Optional[bool] = None) -> DecoderOutput:
    """`Decoder.forward` from precomputed features, `src/milan/decoders.py:379-523`."""
    has_lm = 'lm.embedding.weight' in sd
    if mi is None:
        mi = has_lm and (not isinstance(strategy, str) or strategy != 'rerank')
    n_vocab = len(vocab)
    start_index, stop_index = n_vocab, n_vocab + 1
    batch_size = len(features)
    state = init_state(features, sd, lm=mi)
    currents = torch.full((batch_size,), start_index, dtype=torch.long)
    predictions = attentions = beam_scores = beam_tokens = None
    if isinstance(strategy, torch.Tensor) or strategy == 'greedy':
        V = sd['output.1.weight'].shape[0]
        tokens = currents.new_zeros(batch_size, length)
        scores = features.new_zeros(batch_size)
        predictions = features.new_zeros(batch_size, length, V)
        attentions = features.new_zeros(batch_size, length, features.shape[1])
        for time in range(length):
            outputs = step(features, currents, state, sd, temperature)
            if isinstance(strategy, torch.Tensor):
                currents = strategy[:, time]
            else:
                currents = outputs.predictions.argmax(dim=1)
            predictions[:, time] = outputs.predictions
            attentions[:, time] = outputs.attentions
            tokens[:, time] = currents
            state = outputs.state
            scores += outputs.predictions[torch.arange(batch_size), currents]  # no stop handling (:461-463)
    else:
        runner = BeamSearch(stop_index, max_steps=length, beam_size=beam_size)

        def beam_step(tokens_, st):
            # AllenNLPDecoderState keeps LM states batch-first (src/milan/decoders.py:170-175,192-196).
            h_lm = st['h_lm'].permute(1, 0, 2).contiguous() if 'h_lm' in st else None
            c_lm = st['c_lm'].permute(1, 0, 2).contiguous() if 'c_lm' in st else None
            outputs = step(st['features'], tokens_, DecoderState(st['h'], st['c'], h_lm, c_lm), sd, temperature)
            new = {'features': st['features'], 'h': outputs.state.h, 'c': outputs.state.c}
            if h_lm is not None:
                new['h_lm'] = outputs.state.h_lm.permute(1, 0, 2).contiguous()
                new['c_lm'] = outputs.state.c_lm.permute(1, 0, 2).contiguous()
            return outputs.predictions, new

        start_state = {'features': features, 'h': state.h, 'c': state.c}
        if mi:
            start_state['h_lm'] = state.h_lm.permute(1, 0, 2).contiguous()
            start_state['c_lm'] = state.c_lm.permute(1, 0, 2).contiguous()
        tokens, scores = runner.search(currents, start_state, beam_step)
        beam_scores, beam_tokens = scores, tokens
        if strategy == 'beam':
            tokens, scores = tokens[:, 0], scores[:, 0]
        else:
            starts = currents.new_full((batch_size, beam_size, 1), start_index)
            inputs_lm = torch.cat([starts, tokens], dim=-1).view(batch_size * beam_size, -1)
            scores_lm = lm_forward(inputs_lm, sd, stop_index).view(batch_size, beam_size)
            scores = scores - temperature * scores_lm
            idx_b = torch.arange(batch_size)
            idx_s = scores.argmax(dim=-1)
            tokens = tokens[idx_b, idx_s].view(batch_size, -1)
            scores = scores[idx_b, idx_s].view(batch_size)
    captions = tuple(reconstruct(seq, vocab) for seq in tokens.tolist())
    return DecoderOutput(captions, scores, tokens, predictions, attentions, beam_scores, beam_tokens)


def index_tokens(tokenized: Sequence[Sequence[str]], vocab: Sequence[str], start: bool = True, stop: bool = True,
                 pad: bool = True, unk: bool = True, length: Optional[int] = None) -> Tuple[Tuple[int, ...], ...]:
    """`Indexer.index`, `src/utils/lang.py:460-515`, for an indexer whose defaults are all True / no length."""
    ids = {token: index for index, token in enumerate(vocab)}
    n = len(vocab)
    start_index, stop_index, pad_index, unk_index = n, n + 1, n + 2, n + 3
    length = length or max(len(toks) for toks in tokenized)
    length += int(start) + int(stop)
    indexed = []
    for tokens in tokenized:
        indices = [start_index] if start else []
        if unk:
            indices += [ids.get(tok, unk_index) for tok in tokens]
        else:
            indices += [ids[tok] for tok in tokens if tok in ids]
        if stop:
            if len(indices) >= length:
                indices = indices[:length - 1]
            indices.append(stop_index)
        if len(indices) < length and pad:
            indices += [pad_index] * (length - len(indices))
        elif len(indices) > length:
            indices = indices[:length]
        indexed.append(tuple(indices))
    return tuple(indexed)


@torch.no_grad()
def score(tokenized: Sequence[Sequence[str]], features: torch.Tensor, sd: State, vocab: Sequence[str],
          mi: Optional[bool] = None, temperature: float = 0.2) -> torch.Tensor:
    """`Decoder.score`, `src/milan/decoders.py:636-711`, from pre-tokenized captions and features."""
    if len(features) == 1:
        features = features.expand(len(tokenized), *features.shape[1:])
    targets = torch.tensor(index_tokens(tokenized, vocab))[:, 1:]
    outputs = decode(features, sd, vocab, strategy=targets, length=targets.shape[1], mi=mi, temperature=temperature)
    indexed = index_tokens(tokenized, vocab, start=False, stop=True, pad=False, unk=True)
    totals = [scores[torch.arange(len(indices)), torch.tensor(indices)].sum().item()
              for scores, indices in zip(outputs.predictions, indexed)]
    return torch.tensor(totals)


@torch.no_grad()
def describe(images: torch.Tensor, masks: torch.Tensor, sd: State, vocab: Sequence[str], batch_size: int = 16,
             **kwargs) -> Tuple[str, ...]:
    """`Decoder.predict`, `src/milan/decoders.py:809-871`, for in-memory float exemplars (N,k,3,H,W)/(N,k,1,H,W)."""
    captions: List[str] = []
    for lo in range(0, len(images), batch_size):
        feats = encode(images[lo:lo + batch_size], masks[lo:lo + batch_size], sd)
        captions += decode(feats, sd, vocab, **kwargs).captions
    return tuple(captions)


def to_float_inputs(images_u8: torch.Tensor, masks_u8: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`TopImagesDataset` value contract, `src/milannotations/datasets.py:157,191-197`: images/255, masks float."""
    # Renormalizer('byte' -> 'pt'), src/deps/netdissect/renormalize.py:118-139: data.mul(fp32(1/255)).add_(0).
    scale = torch.tensor(1.0 / 255.0, dtype=torch.float64).to(torch.float32)
    return images_u8.float().mul(scale), masks_u8.float()
