"""Facade of `src/milan/decoders.py`: the reference's `Decoder` surface (constructor, `forward`, `encode`,
`init_state`, `step`, `predict`, checkpoint `load`/`save`) driving the CUDA engine through the C ABI.

Same names, argument meaning and error behaviour as the reference; tensors only cross at the boundary.
There is no CPU path: a `Decoder` must be moved to a CUDA device (`.to('cuda')`) before it can compute.
"""
import collections
from typing import Any, Dict, Mapping, NamedTuple, Optional, Sequence, Tuple, Union

import torch

from neuron_descriptions_b200.engine import Engine
from neuron_descriptions_b200.milan import encoders, lang, lms

StrSequence = Sequence[str]


class DecoderState(NamedTuple):
    """`src/milan/decoders.py:84-99`."""
    h: torch.Tensor
    c: torch.Tensor
    h_lm: Optional[torch.Tensor]
    c_lm: Optional[torch.Tensor]


class DecoderStep(NamedTuple):
    """`src/milan/decoders.py:102-117`."""
    predictions: torch.Tensor
    attentions: torch.Tensor
    state: DecoderState


class DecoderOutput(NamedTuple):
    """`src/milan/decoders.py:120-150`."""
    captions: StrSequence
    scores: torch.Tensor
    tokens: torch.Tensor
    predictions: Optional[torch.Tensor]
    attentions: Optional[torch.Tensor]
    beam_captions: Optional[Sequence[StrSequence]]
    beam_scores: Optional[torch.Tensor]
    beam_tokens: Optional[torch.Tensor]


class _LazyBeamCaptions(collections.abc.Sequence):
    """`beam_captions` (`decoders.py:486-487`) detokenised on first use: 50 strings per neuron are rarely read."""

    def __init__(self, indexer, tokens: torch.Tensor):
        self._indexer, self._tokens, self._value = indexer, tokens, None

    def _materialise(self):
        if self._value is None:
            self._value = tuple(map(self._indexer.reconstruct, self._tokens.tolist()))
        return self._value

    def __len__(self):
        return len(self._tokens)

    def __getitem__(self, index):
        return self._materialise()[index]


Strategy = Union[torch.Tensor, str]
STRATEGY_GREEDY = 'greedy'
STRATEGY_SAMPLE = 'sample'
STRATEGY_BEAM = 'beam'
STRATEGY_RERANK = 'rerank'
STRATEGIES = (STRATEGY_GREEDY, STRATEGY_SAMPLE, STRATEGY_BEAM, STRATEGY_RERANK)


class Decoder:
    """`src/milan/decoders.py:224-1109` (inference surface)."""

    def __init__(self,
                 indexer: lang.Indexer,
                 encoder: encoders.Encoder,
                 lm: Optional[lms.LanguageModel] = None,
                 embedding_size: int = 128,
                 hidden_size: int = 512,
                 attention_hidden_size: Optional[int] = None,
                 dropout: float = .5,
                 length: int = 15,
                 strategy: Optional[str] = None,
                 temperature: float = .2,
                 beam_size: int = 50,
                 precision: str = 'split',
                 max_neurons: int = 64):
        if lm is not None:
            my_vocab, lm_vocab = indexer.vocab.unique, lm.indexer.vocab.unique
            if my_vocab != lm_vocab:
                raise ValueError('lm and decoder have different vocabs;'
                                 f'lm missing {my_vocab - lm_vocab} and decoder missing {lm_vocab - my_vocab}')
        if strategy is None:
            strategy = STRATEGY_BEAM if lm is None else STRATEGY_RERANK
        self.indexer = indexer
        self.encoder = encoder
        self.lm = lm
        self.embedding_size = embedding_size
        self.hidden_size = hidden_size
        self.attention_hidden_size = attention_hidden_size
        self.dropout = dropout
        self.length = length
        self.strategy = strategy
        self.temperature = temperature
        self.beam_size = beam_size
        self.training = False
        self.precision = precision
        self.max_neurons = max_neurons
        self._state_dict: Dict[str, torch.Tensor] = {}
        self._engine: Optional[Engine] = None
        keys_per_image = 49 if getattr(encoder, 'KIND', 'pyramid') == 'spatial' else 1
        self._capacity = {'max_beam': max(50, beam_size), 'max_length': max(15, length),
                          'max_keys': 15 * keys_per_image}

    # ------------------------------------------------------------------ module-ish plumbing
    @property
    def feature_size(self) -> int:
        return self.encoder.feature_shape[-1]

    @property
    def vocab_size(self) -> int:
        return len(self.indexer)

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise RuntimeError('this Decoder is not on a CUDA device: call .to("cuda") first '
                               '(milan_b200 is CUDA-only, there is no CPU fallback)')
        return self._engine

    def state_dict(self) -> Mapping[str, torch.Tensor]:
        return collections.OrderedDict(self._state_dict)

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = False):
        self._state_dict = {k: v.detach().cpu() for k, v in state_dict.items()}
        if self._engine is not None:
            device = self._engine.device
            self._engine.close()
            self._engine = None
            self.to(device)
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError('milan_b200 is an inference engine; training (Decoder.fit) is out of scope')
        self.training = False
        return self

    def to(self, device):
        """Create (or move) the engine. Mirrors `nn.Module.to` for the `predict(device=...)` call path."""
        if device is None:
            return self
        device = torch.device(device)
        if device.type != 'cuda':
            if self._engine is not None:
                self._engine.close()
                self._engine = None
            return self
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is not None and self._engine.device.index == index:
            return self
        if not self._state_dict:
            raise RuntimeError('Decoder has no weights: load a checkpoint or call load_state_dict first')
        if self._engine is not None:
            self._engine.close()
        self._start_staging_allocation(index)  # overlaps the engine build below (weight folding + upload, ~1.5 s)
        self._engine = Engine(self._state_dict, vocab_size=self.vocab_size, device=torch.device('cuda', index),
                              embedding_size=self.embedding_size, hidden_size=self.hidden_size,
                              attention_size=self.attention_hidden_size, feature_size=self.feature_size,
                              lm_embedding_size=self.lm.embedding_size if self.lm else 128,
                              lm_hidden_size=self.lm.hidden_size if self.lm else 512, precision=self.precision,
                              max_neurons=self.max_neurons, encoder_arch=getattr(self.encoder, 'config', 'resnet101'),
                              encoder_kind=getattr(self.encoder, 'KIND', 'pyramid'), **self._capacity)
        if hasattr(self.encoder, 'bind'):
            self.encoder.bind(self._engine)
        if self.lm is not None:
            self.lm.bind(self._engine)
        return self

    def _start_staging_allocation(self, device_index: int, k: int = 15, size: int = 224):
        """Allocate `predict`'s two pinned host slabs (2 x 385 MB at the default capacity; page-locking them takes
        ~0.5 s) on a worker thread while the engine is being built, so that the first exemplar does not wait for it."""
        if getattr(self.encoder, 'KIND', 'pyramid') != 'pyramid' or getattr(self, '_predict_staging', None) is not None \
                or getattr(self, '_staging_future', None) is not None:
            return
        slab = 2 * max(16, (self.max_neurons // 16) * 16)

        def allocate():
            with torch.cuda.device(device_index):
                return [(torch.empty((slab, k, 3, size, size), dtype=torch.uint8, pin_memory=True),
                         torch.empty((slab, k, 1, size, size), dtype=torch.uint8, pin_memory=True)) for _ in range(2)]

        import concurrent.futures
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        self._staging_future = pool.submit(allocate)
        pool.shutdown(wait=False)

    def _ensure_capacity(self, length: int, beam_size: int, n_keys: int):
        """Grow the engine workspace when a call asks for a longer decode, a wider beam or more exemplars per
        neuron than it was sized for (the reference has no such limits)."""
        wanted = {'max_beam': beam_size, 'max_length': length, 'max_keys': n_keys}
        if all(wanted[key] <= self._capacity[key] for key in wanted):
            return
        if wanted['max_beam'] > 64:
            raise ValueError(f'beam_size {beam_size} exceeds the engine limit of 64')
        self._capacity = {key: max(self._capacity[key], wanted[key]) for key in wanted}
        if self._engine is not None:
            device = self._engine.device
            self._engine.close()
            self._engine = None
            self.to(device)

    def cuda(self, device=None):
        return self.to(torch.device('cuda', device) if device is not None else 'cuda')

    def __call__(self, *args, **kwargs) -> DecoderOutput:
        return self.forward(*args, **kwargs)

    # ------------------------------------------------------------------ forward
    def forward(self,
                images_or_features: torch.Tensor,
                masks: Optional[torch.Tensor] = None,
                encode: Optional[bool] = None,
                length: Optional[int] = None,
                strategy: Optional[Strategy] = None,
                mi: Optional[bool] = None,
                temperature: Optional[float] = None,
                beam_size: Optional[int] = None,
                group_size: Optional[int] = None) -> DecoderOutput:
        """`Decoder.forward`, `src/milan/decoders.py:335-523`.

        `group_size` is an extension: when several reference batches are decoded in one call, the reference's
        batch-level early exit (and the LM mask that depends on it) is reproduced per group of that many rows.
        """
        if encode is None:
            encode = masks is not None
        if length is None:
            length = self.length
        if strategy is None:
            strategy = self.strategy
        if mi is None:
            mi = self.lm is not None and not self.training
            mi &= not isinstance(strategy, str) or strategy != STRATEGY_RERANK
        if temperature is None:
            temperature = self.temperature
        if beam_size is None:
            beam_size = self.beam_size
        batch_size = len(images_or_features)

        if mi and isinstance(strategy, str) and strategy == STRATEGY_RERANK:
            raise ValueError('cannot set `mi=` decoding when reranking')
        if (mi or (isinstance(strategy, str) and strategy == STRATEGY_RERANK)) and self.lm is None:
            raise ValueError('cannot use MI/rerank decoding without an LM')
        if (mi or (isinstance(strategy, str) and strategy == STRATEGY_RERANK)) and self.training:
            raise ValueError('cannot use MI/rerank decoding while training')
        if isinstance(strategy, str) and strategy not in STRATEGIES:
            raise ValueError(f'unknown strategy: {strategy}')
        if isinstance(strategy, torch.Tensor):
            if strategy.dim() != 2:
                raise ValueError(f'strategy must be 2D, got {strategy.dim()}')
            if strategy.shape[-1] != length:
                raise ValueError(f'strategy must have length {length}, got {strategy.shape[-1]}')

        self.engine  # raises if not on a CUDA device
        features = self.encode(images_or_features, masks=masks) if encode else images_or_features
        self._ensure_capacity(length, beam_size if strategy in (STRATEGY_BEAM, STRATEGY_RERANK) else 1,
                              features.shape[1])
        engine = self.engine
        features = features.to(engine.device, torch.float32)

        predictions = attentions = beam_captions = beam_scores = beam_tokens = None
        if isinstance(strategy, torch.Tensor) or strategy == STRATEGY_GREEDY:
            forced = strategy if isinstance(strategy, torch.Tensor) else None
            tokens, scores, predictions, attentions = engine.decode_greedy(features, length, mi, temperature, forced)
        elif strategy == STRATEGY_SAMPLE:
            tokens, scores, predictions, attentions = self._sample(features, length, mi, temperature)
        else:
            rerank = strategy == STRATEGY_RERANK
            beam_tokens, beam_scores, steps, tokens, scores, _ = engine.decode_beam(
                features, length, beam_size, rerank, temperature, group_size=group_size or batch_size, mi=mi)
            if group_size is None or group_size >= batch_size:
                # the reference returns only the T <= length columns produced before its early exit
                T = int(steps[0].item())
                beam_tokens, tokens = beam_tokens[..., :T], tokens[..., :T]
            beam_captions = _LazyBeamCaptions(self.indexer, beam_tokens)
        return DecoderOutput(
            captions=self.indexer.reconstruct(tokens.tolist()),
            tokens=tokens,
            scores=scores,
            predictions=predictions,
            attentions=attentions,
            beam_captions=beam_captions,
            beam_scores=beam_scores,
            beam_tokens=beam_tokens,
        )

    def _sample(self, features, length, mi, temperature):
        """strategy='sample' (`decoders.py:448-453`): per-step multinomial draw over the CUDA step's output."""
        batch_size = len(features)
        state = self.init_state(features, lm=mi)
        currents = torch.full((batch_size,), self.indexer.start_index, dtype=torch.long, device=features.device)
        tokens = currents.new_zeros(batch_size, length)
        scores = features.new_zeros(batch_size)
        predictions = features.new_zeros(batch_size, length, self.vocab_size)
        attentions = features.new_zeros(batch_size, length, features.shape[1])
        for time in range(length):
            outputs = self.step(features, currents, state, temperature=temperature)
            currents = torch.multinomial(torch.exp(outputs.predictions), 1).view(batch_size)
            predictions[:, time] = outputs.predictions
            attentions[:, time] = outputs.attentions
            tokens[:, time] = currents
            state = outputs.state
            scores += outputs.predictions.gather(1, currents.view(-1, 1)).view(batch_size)
        return tokens, scores, predictions, attentions

    def encode(self, images: torch.Tensor, masks: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`Decoder.encode`, `src/milan/decoders.py:525-546`."""
        batch_size = len(images)
        images = images.view(-1, *images.shape[-3:])
        if masks is not None:
            masks = masks.view(-1, *masks.shape[-3:])
        features = self.encoder(images, masks=masks)
        return features.view(batch_size, -1, self.feature_size)

    def init_state(self, features: torch.Tensor, lm: bool = True) -> DecoderState:
        """`Decoder.init_state`, `src/milan/decoders.py:548-574`."""
        h, c = self.engine.init_state(features)
        h_lm = c_lm = None
        if self.lm is not None and lm:
            h_lm = h.new_zeros(self.lm.layers, len(features), self.lm.hidden_size)
            c_lm = c.new_zeros(self.lm.layers, len(features), self.lm.hidden_size)
        return DecoderState(h, c, h_lm, c_lm)

    def step(self, features: torch.Tensor, tokens: torch.Tensor, state: DecoderState,
             temperature: Optional[float] = None) -> DecoderStep:
        """`Decoder.step`, `src/milan/decoders.py:576-634`."""
        h, c, h_lm, c_lm = state
        if (h_lm is None) != (c_lm is None):
            raise ValueError('state must have both h_lm and c_lm or neither')
        if h_lm is not None and self.lm is None:
            raise ValueError('state has h_lm or c_lm, but decoder has no lm')
        temperature = self.temperature if temperature is None else temperature
        predictions, attentions, h, c, h_lm, c_lm = self.engine.step(features, tokens, h, c, h_lm, c_lm, temperature)
        return DecoderStep(predictions=predictions, attentions=attentions,
                           state=DecoderState(h=h, c=c, h_lm=h_lm, c_lm=c_lm))

    # ------------------------------------------------------------------ score
    def score(self, captions: StrSequence, images_or_features: torch.Tensor, masks: Optional[torch.Tensor] = None,
              device=None, **kwargs: Any) -> torch.Tensor:
        """`Decoder.score`, `src/milan/decoders.py:636-711`: force-decode the captions and total their
        log-probabilities (`mi=False`) or mutual informations (`mi=True`). Needs an indexer with a tokenizer."""
        for forbidden in ('strategy', 'length'):
            if forbidden in kwargs:
                raise ValueError(f'option disallowed: {forbidden}')
        if masks is not None and len(masks) != len(images_or_features):
            raise ValueError('images_or_features and masks must have the same batch size; '
                             f'got {len(images_or_features)} and {len(masks)}')
        if len(images_or_features) == 1:
            images_or_features = images_or_features.expand(len(captions), *images_or_features.shape[1:])
            if masks is not None:
                masks = masks.expand(len(captions), *masks.shape[1:])
        elif len(images_or_features) != len(captions):
            raise ValueError('images_or_features must have batch size 1 or '
                             f'{len(captions)}; got {len(images_or_features)}')
        if device is not None:
            self.to(device)
        targets = torch.tensor(self.indexer(captions))
        targets = targets[:, 1:]
        _, length = targets.shape
        totals = []
        indexed = self.indexer(captions, start=False, stop=True, pad=False, unk=True)
        chunk = self.engine.cfg.max_neurons
        for lo in range(0, len(captions), chunk):
            hi = min(lo + chunk, len(captions))
            outputs = self(images_or_features[lo:hi], masks=None if masks is None else masks[lo:hi],
                           strategy=targets[lo:hi], length=length, **kwargs)
            predictions = outputs.predictions.cpu()
            for scores, indices in zip(predictions, indexed[lo:hi]):
                totals.append(scores[torch.arange(len(indices)), torch.tensor(indices, dtype=torch.long)].sum().item())
        return torch.tensor(totals, device=self.engine.device if device is not None else None)

    # ------------------------------------------------------------------ predict
    def predict(self,
                dataset,
                mask: bool = True,
                image_index: int = 2,
                mask_index: int = 3,
                batch_size: int = 16,
                features=None,
                num_workers: int = 0,
                device=None,
                display_progress_as: Optional[str] = 'predict captions',
                **kwargs: Any) -> StrSequence:
        """`Decoder.predict`, `src/milan/decoders.py:809-871`.

        Several reference batches are fused into one engine call (`group_size=batch_size` keeps the reference's
        per-batch early-exit semantics); datasets exposing `batch_u8(lo, hi)` are fed as uint8 (4x less H2D).
        """
        del num_workers
        if device is not None:
            self.to(device)
        engine = self.engine
        fast = self._predict_host_pipeline(dataset, mask, batch_size, features, display_progress_as, kwargs)
        if fast is not None:
            return fast
        source = dataset if features is None else features
        total = len(source)
        chunk = max(batch_size, (engine.cfg.max_neurons // batch_size) * batch_size)
        chunk = min(chunk, engine.cfg.max_neurons) if engine.cfg.max_neurons < batch_size else chunk
        captions = []
        token_rows = []
        length = kwargs.get('length') or self.length
        bounds = [(lo, min(lo + chunk, total)) for lo in range(0, total, chunk)]

        # Host feed: uint8 batches are assembled from the mmapped exemplar files into two alternating pinned staging
        # buffers by a worker thread while the GPU works on the previous chunk (numpy copies release the GIL).
        use_u8 = features is None and hasattr(dataset, 'batch_u8')
        pool = staging = None
        if use_u8:
            import concurrent.futures
            pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
            staging = [dataset.alloc_batch_u8(chunk) if hasattr(dataset, 'alloc_batch_u8') else None for _ in range(2)]

            def load(i):
                lo_, hi_ = bounds[i]
                if staging[i % 2] is not None:
                    return dataset.batch_u8(lo_, hi_, out=staging[i % 2])
                return dataset.batch_u8(lo_, hi_)

            pending = pool.submit(load, 0) if bounds else None
        iterator = range(len(bounds))
        if display_progress_as is not None:
            try:
                from tqdm.auto import tqdm
                iterator = tqdm(iterator, desc=display_progress_as)
            except ImportError:
                pass
        for i in iterator:
            lo, hi = bounds[i]
            group = batch_size if chunk > batch_size else None
            with torch.no_grad():
                if features is not None:
                    feats = torch.stack([torch.as_tensor(features[j][0]) for j in range(lo, hi)])
                    output = self(feats, group_size=group, **kwargs)
                else:
                    if use_u8:
                        images, masks = pending.result()
                        images = images.to(engine.device, non_blocking=True)
                        masks = masks.to(engine.device, non_blocking=True) if mask else None
                        torch.cuda.current_stream(engine.device).synchronize()  # staging buffer is free again
                        if i + 1 < len(bounds):
                            pending = pool.submit(load, i + 1)
                    else:
                        samples = [dataset[j] for j in range(lo, hi)]
                        images = torch.stack([torch.as_tensor(s[image_index]) for s in samples])
                        images = images.to(engine.device, non_blocking=True)
                        masks = None
                        if mask:  # the mask field is only read when asked for (decoders.py:853-858)
                            masks = torch.stack([torch.as_tensor(s[mask_index]) for s in samples])
                            masks = masks.to(engine.device, non_blocking=True)
                    if masks is None:
                        output = self(images, encode=True, group_size=group, **kwargs)
                    else:
                        output = self(images, masks, group_size=group, **kwargs)
            captions += output.captions
            rows = torch.full((len(output.tokens), length), self.indexer.stop_index, dtype=torch.long)
            rows[:, :output.tokens.shape[1]] = output.tokens.cpu()
            token_rows.append(rows)
        if pool is not None:
            pool.shutdown(wait=True)
        self.last_predict_tokens = torch.cat(token_rows) if token_rows else torch.empty(0, length, dtype=torch.long)
        return tuple(captions)

    def _predict_host_pipeline(self, dataset, mask, batch_size, features, display_progress_as, kwargs):
        """`predict` for the common case — a uint8 exemplar dataset (`batch_u8`), masks on, a string strategy — as
        slabs of several engine chunks through `milan_describe_host`, whose copy stream moves the exemplars of
        chunk i+1 while chunk i is encoded / decoded; a worker thread fills the next pinned slab from the mmapped
        files meanwhile. Returns None when the call needs the general path."""
        allowed = {'strategy', 'temperature', 'beam_size', 'length', 'mi'}
        if features is not None or not mask or not hasattr(dataset, 'batch_u8') or set(kwargs) - allowed:
            return None
        if getattr(dataset, 'transform_images', None) is not None or getattr(dataset, 'transform_masks', None) is not None:
            return None
        strategy = kwargs.get('strategy') or self.strategy
        if not isinstance(strategy, str) or strategy not in (STRATEGY_GREEDY, STRATEGY_BEAM, STRATEGY_RERANK):
            return None
        engine = self.engine
        if engine.cfg.max_neurons < batch_size or engine.keys_per_image != 1 or len(dataset) == 0:
            return None
        length = kwargs.get('length') or self.length
        beam_size = kwargs.get('beam_size') or self.beam_size
        temperature = self.temperature if kwargs.get('temperature') is None else kwargs['temperature']
        mi = kwargs.get('mi')
        if mi is None:
            mi = self.lm is not None and strategy != STRATEGY_RERANK
        if mi and strategy == STRATEGY_RERANK:
            raise ValueError('cannot set `mi=` decoding when reranking')
        if (mi or strategy == STRATEGY_RERANK) and self.lm is None:
            raise ValueError('cannot use MI/rerank decoding without an LM')
        import time
        clock = time.perf_counter
        t0 = clock()
        self._ensure_capacity(length, beam_size if strategy != STRATEGY_GREEDY else 1, getattr(dataset, 'k', 15))
        engine = self.engine
        t_capacity = clock()
        import concurrent.futures
        total = len(dataset)
        chunk = (engine.cfg.max_neurons // batch_size) * batch_size
        slab = 2 * chunk
        bounds = [(lo, min(lo + slab, total)) for lo in range(0, total, slab)]
        # pinned staging slabs are expensive to allocate: keep them on the decoder between calls
        cached = getattr(self, '_predict_staging', None)
        future = getattr(self, '_staging_future', None)
        if cached is None and future is not None:
            self._staging_future = None
            try:
                cached = future.result()
            except RuntimeError:  # pinned allocation failed in the background: fall through to a fresh (smaller) one
                cached = None
        want = min(slab, total)
        if cached is None or cached[0][0].shape[0] < want or cached[0][0].shape[1:] != dataset.alloc_batch_u8(0)[0].shape[1:]:
            cached = [dataset.alloc_batch_u8(want) for _ in range(2)]
            self._predict_staging = cached
        staging = cached
        t_staging = clock()
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        wait_s, call_s = [], []

        def load(i):
            lo, hi = bounds[i]
            return dataset.batch_u8(lo, hi, out=staging[i % len(staging)])

        pending = pool.submit(load, 0)
        iterator = range(len(bounds))
        if display_progress_as is not None:
            try:
                from tqdm.auto import tqdm
                iterator = tqdm(iterator, desc=display_progress_as)
            except ImportError:
                pass
        token_rows = []
        for i in iterator:
            t_wait = clock()
            images, masks = pending.result()
            if i + 1 < len(bounds):
                pending = pool.submit(load, i + 1)
            t_call = clock()
            with torch.no_grad():
                tokens, _, steps = engine.describe_host(images, masks, strategy=strategy, mi=mi, length=length,
                                                        beam=beam_size, group_size=batch_size,
                                                        temperature=temperature)
            wait_s.append(t_call - t_wait)
            call_s.append(clock() - t_call)
            if strategy != STRATEGY_GREEDY:
                # the reference returns only the T <= length columns decoded before its per-batch early exit; pad
                # the rest with <stop> (reconstruct cuts at the first <stop> anyway)
                keep = torch.arange(length).unsqueeze(0) < steps.to(torch.long).unsqueeze(1)
                tokens = torch.where(keep, tokens, torch.full_like(tokens, self.indexer.stop_index))
            token_rows.append(tokens)
        pool.shutdown(wait=True)
        self.last_predict_tokens = torch.cat(token_rows)
        t_described = clock()
        captions = tuple(self.indexer.reconstruct(self.last_predict_tokens.tolist()))
        # where the wall-clock of this call went (seconds): read by scripts/compute_milan_descriptions.py's timing dump
        self.last_predict_timing = {
            'neurons': total, 'slabs': len(bounds), 'ensure_capacity_s': t_capacity - t0,
            'staging_alloc_s': t_staging - t_capacity, 'feed_wait_s': sum(wait_s), 'first_feed_wait_s': wait_s[0],
            'describe_calls_s': sum(call_s), 'first_describe_call_s': call_s[0], 'detokenise_s': clock() - t_described}
        return captions

    def fit(self, *args, **kwargs):
        raise NotImplementedError('training (src/milan/decoders.py:873-1070) is out of scope for this engine')

    # ------------------------------------------------------------------ checkpoints
    def properties(self) -> Mapping[str, Any]:
        """`src/milan/decoders.py:1072-1086`."""
        return {
            'indexer': self.indexer, 'encoder': self.encoder, 'lm': self.lm,
            'embedding_size': self.embedding_size, 'hidden_size': self.hidden_size,
            'attention_hidden_size': self.attention_hidden_size, 'dropout': self.dropout, 'length': self.length,
            'strategy': self.strategy, 'temperature': self.temperature, 'beam_size': self.beam_size,
        }

    def serialize(self) -> Mapping[str, Any]:
        """Reference payload layout (`src/utils/serialize.py:80-118,188-203`). A tokenizer that came in with a
        reference checkpoint is written back as the opaque payload it arrived as; callables are not serialized."""
        def indexer_payload(indexer):
            props = {'vocab': {'properties': {'tokens': tuple(indexer.vocab.tokens)}, 'children': {}},
                     'tokenize': indexer.tokenize_payload, 'start': indexer.start, 'stop': indexer.stop, 'pad': indexer.pad,
                     'unk': indexer.unk, 'length': indexer.length}
            return {'properties': props, 'children': {}}

        props = dict(self.properties())
        props['indexer'] = indexer_payload(self.indexer)
        props['encoder'] = {'properties': dict(self.encoder.properties()), 'children': {}}
        if self.lm is not None:
            lm_props = dict(self.lm.properties())
            lm_props['indexer'] = indexer_payload(self.lm.indexer)
            props['lm'] = {'properties': lm_props, 'children': {}}
        return {'properties': props, 'children': {'encoder': encoders.key(self.encoder)},
                'state_dict': self.state_dict()}

    def save(self, file, **kwargs: Any) -> None:
        torch.save(self.serialize(), file, **kwargs)

    @classmethod
    def deserialize(cls, payload: Mapping[str, Any], **overrides: Any) -> 'Decoder':
        """Rebuild from a reference checkpoint payload (`serialize.py:121-163,221-253`; `decoders.py:1095-1109`)
        without spaCy or a network: the tokenizer is dropped and `pretrained=False` is implied."""
        props = dict(payload['properties'])
        children = dict(payload.get('children', {}))
        encoder_key = children.get('encoder')
        if encoder_key is None:
            raise ValueError('serialized decoder missing encoder')
        indexer = lang.indexer_from_payload(props['indexer'])
        encoder_props = dict(props['encoder']['properties'])
        encoder = encoders.parse(encoder_key)(**encoder_props)
        lm = None
        if props.get('lm') is not None:
            lm_props = dict(props['lm']['properties'])
            lm_indexer = lang.indexer_from_payload(lm_props.pop('indexer'))
            lm = lms.LanguageModel(lm_indexer, **lm_props)
        kwargs = {key: props[key] for key in ('embedding_size', 'hidden_size', 'attention_hidden_size', 'dropout',
                                              'length', 'strategy', 'temperature', 'beam_size') if key in props}
        if 'reranker_kwargs' in props and hasattr(cls, 'from_decoder'):  # a DecoderWithCLIP payload (`:1198-1203`)
            kwargs['reranker_kwargs'] = props['reranker_kwargs']
        kwargs.update(overrides)
        decoder = cls(indexer, encoder, lm=lm, **kwargs)
        state_dict = payload.get('state_dict')
        if state_dict is not None:
            decoder.load_state_dict(state_dict, strict=False)
        return decoder.eval()

    @classmethod
    def load(cls, file, **kwargs: Any) -> 'Decoder':
        """`SerializableModule.load`, `src/utils/serialize.py:255-269`; kwargs go to `torch.load`."""
        overrides = {key: kwargs.pop(key) for key in ('precision', 'max_neurons', 'reranker', 'reranker_kwargs')
                     if key in kwargs}
        kwargs.setdefault('map_location', 'cpu')
        kwargs.setdefault('weights_only', False)
        payload = torch.load(file, **kwargs)
        return cls.deserialize(payload, **overrides)


ENGINE_MAX_BEAM = 64  # the beam kernels rank one source row per warp of an 8-CTA cluster (csrc/decode_fused.h)


class DecoderWithCLIP(Decoder):
    """`DecoderWithCLIP`, `src/milan/decoders.py:1115-1211`: beam-decode with MILAN, then let an image-text similarity
    model rerank each neuron's beam against its masked and unmasked exemplars.

    Differences from the reference, both forced by what is available offline:
      * the similarity model is injected (`reranker=` a `rerankers.SimilarityReranker`, or `reranker_kwargs` with a
        `similarity` callable) instead of being CLIP ViT-B/32 loaded by name (`rerankers.reranker`);
      * the default beam is the engine's widest, 64, where the reference samples 1000 candidates (`:1124-1126`).
    """

    def __init__(self, *args: Any, reranker=None, reranker_kwargs: Optional[Mapping[str, Any]] = None, **kwargs: Any):
        from neuron_descriptions_b200.milan import rerankers
        kwargs.setdefault('strategy', STRATEGY_BEAM)
        kwargs.setdefault('beam_size', ENGINE_MAX_BEAM)
        kwargs.setdefault('temperature', .5)
        super().__init__(*args, **kwargs)
        self.reranker_kwargs = dict(reranker_kwargs) if reranker_kwargs else {}
        self.reranker = reranker if reranker is not None else rerankers.reranker(**self.reranker_kwargs)

    def forward(self, images_or_features: torch.Tensor, masks: Optional[torch.Tensor] = None,  # type: ignore[override]
                lam: Optional[float] = None, **kwargs: Any) -> DecoderOutput:
        """Captions / scores / tokens are those of the beam entry the reranker puts first; the remaining fields are
        the beam decode's (`:1135-1196`). The images must be real images: the reranker sees them too."""
        if masks is None:
            raise ValueError('must specify masks in DecoderWithCLIP')
        if 'strategy' in kwargs:
            raise ValueError('cannot set "strategy" in DecoderWithCLIP')
        images = images_or_features
        outputs = super().forward(images, masks=masks, strategy=STRATEGY_BEAM, **kwargs)
        assert outputs.beam_captions is not None and outputs.beam_scores is not None and outputs.beam_tokens is not None
        reranked = self.reranker(images, masks, outputs.beam_captions, lam=lam)
        first = torch.tensor([order[0] for order in reranked.orders], device=outputs.beam_scores.device)
        rows = torch.arange(len(first), device=first.device)
        return DecoderOutput(tuple(texts[0] for texts in reranked.texts), outputs.beam_scores[rows, first],
                             outputs.beam_tokens[rows, first], *outputs[3:])

    def properties(self) -> Mapping[str, Any]:
        """`:1198-1203`. (A similarity callable inside `reranker_kwargs` is not serializable and is left out.)"""
        kept = {key: value for key, value in self.reranker_kwargs.items() if not callable(value)}
        return {**super().properties(), 'reranker_kwargs': kept}

    @classmethod
    def from_decoder(cls, decoder: Decoder, **kwargs: Any) -> 'DecoderWithCLIP':
        """`:1205-1209`: a base `Decoder` with a reranker on top; `kwargs` carry `reranker=` / `reranker_kwargs=`."""
        return cls.deserialize(decoder.serialize(), **kwargs)


def decoder(*args, **kwargs):
    raise NotImplementedError('Decoder training factory (src/milan/decoders.py:1214) is out of scope')
