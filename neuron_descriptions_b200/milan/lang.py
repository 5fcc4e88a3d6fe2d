"""Host-side vocabulary / detokenisation mirror of `src/utils/lang.py` (only what decoding needs).

`Indexer.reconstruct` / `unindex` follow `src/utils/lang.py:573-612,678-730`; special ids follow `:242-260`;
`Indexer.index` / `__call__` follow `:393-515` (needed by `Decoder.score`). The reference tokenises with spaCy
(`lang.Tokenizer`, `:14-71`), which is not available offline and whose pipeline bytes inside real checkpoints
cannot be rebuilt without it: `tokenize` is therefore any callable `Sequence[str] -> Sequence[Sequence[str]]`;
`BasicTokenizer` is a dependency-free stand-in (lower-cased word split, punctuation dropped, NO lemmatisation
and NO stop-word removal -- it is an approximation of the spaCy pipeline, not a replacement). An `Indexer`
built with `tokenize=None` (every checkpoint loaded here) raises if asked to index raw text.
"""
import re
import collections
import dataclasses
import functools
from typing import Any, Callable, Mapping, Optional, Sequence, Tuple, Union

START_TOKEN = '<start>'
STOP_TOKEN = '<stop>'
PAD_TOKEN = '<pad>'
UNK_TOKEN = '<unk>'


@dataclasses.dataclass(frozen=True)
class Vocab:
    """`src/utils/lang.py:93-178`."""

    tokens: Tuple[str, ...]

    def __post_init__(self):
        object.__setattr__(self, 'tokens', tuple(self.tokens))

    @functools.cached_property
    def ids(self) -> Mapping[str, int]:
        return {token: index for index, token in enumerate(self.tokens)}

    @functools.cached_property
    def unique(self):
        return frozenset(self.ids)

    def __getitem__(self, token):
        if isinstance(token, (int, slice)):
            return self.tokens[token]
        return self.ids[token]

    def __len__(self) -> int:
        return len(self.tokens)

    def __contains__(self, token) -> bool:
        if isinstance(token, int):
            return 0 <= token < len(self)
        return token in self.unique

    def properties(self):
        return {'tokens': self.tokens}


@dataclasses.dataclass(frozen=True)
class Indexer:
    """`src/utils/lang.py:230-747`: ids <-> text with the four special tokens after the vocabulary."""

    vocab: Vocab
    tokenize: Optional[Callable[..., Any]] = None
    start: bool = False
    stop: bool = False
    pad: bool = False
    unk: bool = False
    length: Optional[int] = None
    # A reference checkpoint's serialized tokenizer ({'properties': {'nlp': (spaCy config, bytes), ...}, 'children':
    # {}}, src/utils/serialize.py:104-107): opaque here (no spaCy), carried so that `Decoder.save` round-trips it.
    tokenize_payload: Optional[Any] = dataclasses.field(default=None, compare=False, repr=False)

    @functools.cached_property
    def start_index(self) -> int:
        return len(self.vocab)

    @functools.cached_property
    def stop_index(self) -> int:
        return len(self.vocab) + 1

    @functools.cached_property
    def pad_index(self) -> int:
        return len(self.vocab) + 2

    @functools.cached_property
    def unk_index(self) -> int:
        return len(self.vocab) + 3

    @functools.cached_property
    def specials(self) -> Mapping[int, str]:
        return collections.OrderedDict((
            (self.start_index, START_TOKEN),
            (self.stop_index, STOP_TOKEN),
            (self.pad_index, PAD_TOKEN),
            (self.unk_index, UNK_TOKEN),
        ))

    @functools.cached_property
    def tokens(self) -> Tuple[str, ...]:
        return tuple(list(self.vocab.tokens) + list(self.specials.values()))

    @functools.cached_property
    def ids(self) -> Mapping[str, int]:
        ids = dict(self.vocab.ids)
        for index, token in self.specials.items():
            ids[token] = index
        return ids

    @functools.cached_property
    def unique(self):
        return frozenset(self.ids)

    def __getitem__(self, token):
        if isinstance(token, (int, slice)):
            return self.tokens[token]
        return self.ids[token]

    def __len__(self) -> int:
        return len(self.vocab) + len(self.specials)

    def __contains__(self, token) -> bool:
        if isinstance(token, int):
            return 0 <= token < len(self)
        return token in self.unique

    def __call__(self, texts, **kwargs):
        """`Indexer.__call__`, `src/utils/lang.py:379-391`: tokenize then `index`."""
        if self.tokenize is None:
            raise NotImplementedError(
                'this Indexer has no tokenizer (checkpoints are loaded without the reference\'s spaCy pipeline, '
                'src/utils/lang.py:14-71); pass tokenize=<callable> (e.g. lang.BasicTokenizer()) or call '
                '.index() with pre-tokenized text')
        tokenized = self.tokenize([texts] if isinstance(texts, str) else texts)
        indexed = self.index(tokenized, **kwargs)
        return indexed[0] if isinstance(texts, str) else indexed

    def index(self, tokenized, start: Optional[bool] = None, stop: Optional[bool] = None,
              pad: Optional[bool] = None, unk: Optional[bool] = None, length: Optional[int] = None):
        """`Indexer.index`, `src/utils/lang.py:460-515`: tokens -> ids.

        Same observable behaviour as the reference, including its corner cases: `length` counts vocabulary tokens
        only (one slot is added per enabled start / stop token); without `length` the longest input decides (for a
        single flat sequence that is the longest TOKEN STRING, as in the reference); with `stop` the sequence is
        truncated so that `<stop>` always fits; unknown tokens become `<unk>` or are dropped.
        """
        if not tokenized:
            return ()
        flat = isinstance(tokenized[0], str)
        use = {name: (getattr(self, name) if flag is None else flag)
               for name, flag in (('start', start), ('stop', stop), ('pad', pad), ('unk', unk))}
        budget = (length or self.length or max(map(len, tokenized))) + int(use['start']) + int(use['stop'])
        lookup = self.vocab.ids

        def encode(tokens):
            ids = [self.start_index] if use['start'] else []
            if use['unk']:
                ids.extend(lookup.get(token, self.unk_index) for token in tokens)
            else:
                ids.extend(lookup[token] for token in tokens if token in lookup)
            if use['stop']:
                del ids[max(budget - 1, 0):]
                ids.append(self.stop_index)
            if use['pad'] and len(ids) < budget:
                ids.extend([self.pad_index] * (budget - len(ids)))
            return tuple(ids[:budget])

        if flat:
            return encode(tokenized)
        return tuple(encode(tokens) for tokens in tokenized)

    def unindex(self, indexed, specials: bool = True, start: bool = True, stop: bool = True, pad: bool = True,
                unk: bool = True):
        """ids -> token strings (`Indexer.unindex`, `src/utils/lang.py:573-612`).

        Vocabulary ids map to their token; a special id maps to its token when `specials` and its own flag are set
        and is dropped otherwise; anything else raises ValueError. As in the reference all four special ids are
        recognised whether or not the indexer itself emits them.
        """
        if not indexed:
            return ()
        special_text = {index: (token if specials and wanted else None)
                        for (index, token), wanted in zip(self.specials.items(), (start, stop, pad, unk))}
        n_vocab = len(self.vocab)

        def decode(indices):
            tokens = []
            for index in indices:
                if index < n_vocab:
                    tokens.append(self.vocab[index])
                elif index in special_text:
                    if special_text[index] is not None:
                        tokens.append(special_text[index])
                else:
                    raise ValueError(f'unknown index: {index}')
            return tuple(tokens)

        if isinstance(indexed[0], int):
            return decode(indexed)
        return tuple(decode(indices) for indices in indexed)

    @staticmethod
    def _detokenize(tokens, dropped) -> str:
        """Words -> caption text with the reference's surface rules (`src/utils/lang.py:704-728`): cut at the
        first `<stop>`, drop special tokens, no space before . , ; : and none around -, sentences capitalised."""
        tokens = list(tokens)
        if STOP_TOKEN in tokens:
            del tokens[tokens.index(STOP_TOKEN):]
        text = ' '.join(token for token in tokens if token not in dropped)
        for mark in '.,;:':
            text = text.replace(' ' + mark, mark)
        text = text.replace(' -', '-').replace('- ', '-')
        sentences = (sentence.strip().capitalize() for sentence in text.split('.'))
        return '. '.join(sentences).strip()

    def reconstruct(self, inputs: Union[Sequence[int], Sequence[Sequence[int]], Sequence[str],
                                        Sequence[Sequence[str]]]):
        """ids or tokens -> caption strings (`Indexer.reconstruct`, `src/utils/lang.py:678-730`); one sequence in,
        one string out, several in, a tuple out."""
        if not inputs:
            raise ValueError('must provide at least one seq')
        for position, item in enumerate(inputs):
            if not isinstance(item, (int, str)) and not item:
                raise ValueError(f'input seq {position} is empty')
        single = isinstance(inputs[0], (int, str))
        batch = [inputs] if single else inputs
        probe = batch[0][0]
        assert isinstance(probe, (int, str)), 'unknown input type'
        if isinstance(probe, int):
            batch = self.unindex(batch)
        dropped = frozenset(self.specials.values())
        texts = tuple(self._detokenize(tokens, dropped) for tokens in batch)
        return texts[0] if single else texts

    def properties(self):
        return {'vocab': self.vocab, 'tokenize': self.tokenize, 'start': self.start, 'stop': self.stop,
                'pad': self.pad, 'unk': self.unk, 'length': self.length}


class BasicTokenizer:
    """Dependency-free stand-in for `lang.Tokenizer` (`src/utils/lang.py:14-71`): lower-cased word / number
    tokens, punctuation dropped. No lemmatisation or stop-word list (those need spaCy's `en_core_web_sm`)."""

    _WORD = re.compile(r"[A-Za-z0-9]+(?:'[a-z]+)?")

    def __init__(self, lowercase: bool = True):
        self.lowercase = lowercase

    def __call__(self, texts):
        singleton = isinstance(texts, str)
        out = []
        for text in [texts] if singleton else texts:
            tokens = self._WORD.findall(text)
            out.append(tuple(tok.lower() if self.lowercase else tok for tok in tokens))
        return out[0] if singleton else tuple(out)


def indexer_from_payload(payload: Mapping[str, Any]) -> Indexer:
    """Rebuild an `Indexer` from a reference `Serializable` payload (`src/utils/serialize.py:121-163`) without
    spaCy: the tokenizer child is kept as its raw payload (decoding never tokenises)."""
    props = dict(payload['properties'])
    vocab = props['vocab']
    if isinstance(vocab, Mapping):
        vocab = Vocab(tuple(vocab['properties']['tokens']))
    return Indexer(vocab=vocab, tokenize=None, start=bool(props.get('start', False)),
                   stop=bool(props.get('stop', False)), pad=bool(props.get('pad', False)),
                   unk=bool(props.get('unk', False)), length=props.get('length'),
                   tokenize_payload=props.get('tokenize'))
