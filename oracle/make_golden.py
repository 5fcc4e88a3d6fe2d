"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on CPU.

Run in the build container only (`python -m oracle.make_golden`); the GPU box has no /root/reference.
Weights and inputs are regenerated from seeds by `neuron_descriptions_b200.synthetic`, so only outputs are
stored. The beam search inside the reference's `Decoder.forward` is `oracle.beam_search.BeamSearch`
(allennlp is not installed; see that file's header) — every other line executed is the reference's own.
"""
import os
import pathlib
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neuron_descriptions_b200 import synthetic  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
VARIANTS = {  # name -> (sharpen, stop_bias)
    'flat': (1.0, 0.0),     # default nn init: near-uniform logits, top-k order is rounding-sensitive
    'sharp': (12.0, 0.0),   # well separated, full-length beams
    'stop': (12.0, 2.0),    # beams end with <stop> at varied lengths (forced-stop path), T = 15
    'early': (8.0, 2.0),    # every beam ends early -> early exit, T < 15
}
ENC_NEURONS, DEC_NEURONS, K = 2, 6, 15


def whitespace_tokenize(texts):
    """Stand-in tokenizer for the score golden (spaCy is not installed): any callable works as `Indexer.tokenize`."""
    return tuple(tuple(text.lower().split()) for text in texts)


SCORE_CAPTIONS = ('the dog and w4999', 'cat', 'w100 zzz sky sky grass w2048 w77', 'Red blue ROUND text faces',
                  'w31 w32 w33 w34 w35 w36 w37 w38 w39 w40 w41 w42', 'unknownword')


def build_reference_decoder(milan, lang, sd, vocab, with_encoder=True, tokenize=None):
    indexer = lang.Indexer(lang.Vocab(tuple(vocab)), tokenize=tokenize, start=True, stop=True, pad=True, unk=True)
    encoder = milan.encoders.PyramidConvEncoder('resnet101', pretrained=False)
    lm = milan.lms.LanguageModel(indexer)
    decoder = milan.decoders.Decoder(indexer, encoder, lm=lm)
    missing, unexpected = decoder.load_state_dict(sd, strict=False)
    if with_encoder:
        assert not missing and not unexpected, (missing, unexpected)
    else:
        assert all(k.startswith('encoder.') for k in missing) and not unexpected, (missing, unexpected)
    return decoder.eval()


def synthetic_features(n, k, seed):
    """Feature tensors with the sparsity/magnitude of masked-pooled ResNet activations."""
    gen = torch.Generator().manual_seed(seed + 4242)
    feats = torch.randn(n, k, synthetic.FEATURE_SIZE, generator=gen).abs() * 0.5
    feats[:, :, :64] = torch.randn(n, k, 64, generator=gen) * 0.3  # raw conv1 block can be negative
    feats[0, 1] = 0.0  # an all-zero-mask exemplar
    return feats


def make_score_golden(milan, lang, vocab):
    """`Decoder.score` (src/milan/decoders.py:636-711) + `Indexer.__call__/index` (src/utils/lang.py:379-515)."""
    feats = synthetic_features(DEC_NEURONS, K, seed=0)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0, with_encoder=False)
    decoder = build_reference_decoder(milan, lang, sd, vocab, with_encoder=False, tokenize=whitespace_tokenize)
    captions = list(SCORE_CAPTIONS)
    with torch.no_grad():
        scores = decoder.score(captions, feats, mi=False)
        scores_mi = decoder.score(captions, feats)  # mi defaults to True with an LM (decoders.py:385-387)
        scores_one = decoder.score(captions, feats[:1], mi=False)  # one feature set for every caption
    indexer = decoder.indexer
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, 'score.npz'),
        indexed_default=np.array(indexer(captions), dtype=np.int64),
        indexed_nopad=np.array([list(row) + [-1] * (16 - len(row))
                                for row in indexer(captions, start=False, stop=True, pad=False, unk=True)],
                               dtype=np.int64),
        indexed_len4=np.array(indexer(captions, length=4), dtype=np.int64),
        indexed_nounk=np.array([list(row) + [-1] * (16 - len(row))
                                for row in indexer(captions, start=False, stop=False, pad=False, unk=False)],
                               dtype=np.int64),
        indexed_single=np.array(indexer(captions[2]), dtype=np.int64),
        scores=scores.numpy(), scores_mi=scores_mi.numpy(), scores_one=scores_one.numpy())
    print('score golden:', scores.tolist(), scores_mi.tolist())


ENCODER_VARIANTS = (('pyramid', 'resnet18'), ('pyramid', 'resnet50'), ('spatial', 'resnet18'), ('pyramid', 'alexnet'))


def encoder_variant_inputs():
    """5 exemplar images + masks for the secondary encoder configs (one all-zero mask, one tiny mask)."""
    images_u8, masks_u8 = synthetic.synthetic_exemplars(1, 5, seed=5, zero_mask_fraction=0.0)
    masks_u8[0, 1] = 0
    masks_u8[0, 2] = 0
    masks_u8[0, 2, :, 100:102, 50:52] = 1
    return images_u8.view(5, 3, 224, 224), masks_u8.view(5, 1, 224, 224)


def make_encoder_variant_goldens(milan):
    """The reference's `PyramidConvEncoder('resnet18'|'resnet50')` and `SpatialConvEncoder('resnet18')`
    (`src/milan/encoders.py:159-351`) on seeded weights and exemplars."""
    images_u8, masks_u8 = encoder_variant_inputs()
    scale = torch.tensor(1.0 / 255.0, dtype=torch.float64).to(torch.float32)
    images, masks = images_u8.float().mul(scale), masks_u8.float()
    out = {}
    for kind, arch in ENCODER_VARIANTS:
        cls = milan.encoders.SpatialConvEncoder if kind == 'spatial' else milan.encoders.PyramidConvEncoder
        encoder = cls(arch, pretrained=False)
        missing, unexpected = encoder.load_state_dict(synthetic.synthetic_encoder_state_dict(arch, seed=3), strict=False)
        assert not unexpected and all('classifier' in key for key in missing), (missing, unexpected)
        encoder.eval()
        with torch.no_grad():
            features = encoder(images, masks)
        assert tuple(features.shape[1:]) == tuple(encoder.feature_shape)
        out[f'{kind}_{arch}'] = features.numpy()
        print(f'encoder golden [{kind}/{arch}]:', tuple(features.shape), 'abs max', features.abs().max().item())
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'encoder_variants.npz'), **out)


def exemplar_toy_model(seed: int = 0):
    """The toy CNN of the reference's own stage-1 tests (`tests/exemplars/compute_test.py:149-165`): two 4x4 convs."""
    import collections
    gen = torch.Generator().manual_seed(seed)
    model = torch.nn.Sequential(collections.OrderedDict([
        ('conv_1', torch.nn.Conv2d(3, 6, 4, padding=2)), ('conv_2', torch.nn.Conv2d(6, 6, 4, padding=2))]))
    with torch.no_grad():
        for param in model.parameters():
            param.copy_(torch.randn(param.shape, generator=gen) * 0.3)
    return model.eval()


def exemplar_toy_images(n: int = 20, size: int = 16, seed: int = 0):
    return torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(seed + 5))


EXEMPLAR_CASES = (('conv_1', 16, 4), ('conv_2', 24, 3))  # (layer, output_size, k)


def make_exemplars_golden():
    """The reference's `exemplars.compute.discriminative` (`src/exemplars/compute.py:263-349`) on the toy model:
    20 images x 17x17 (18x18) positions stay below the 8192-sample capacity of the quantile sketch's first level,
    where the sketch is exact and deterministic."""
    import tempfile
    from torch.utils import data
    compute = ref_import.import_reference_exemplars()
    model, images = exemplar_toy_model(), exemplar_toy_images()
    out = {}
    for layer, output_size, k in EXEMPLAR_CASES:
        root = pathlib.Path(tempfile.mkdtemp())
        compute.discriminative(model, data.TensorDataset(images), layer=layer, device='cpu', results_dir=root / 'res',
                               viz_dir=root / 'viz', display_progress=False, num_workers=0, k=k, quantile=0.99,
                               image_size=16, output_size=output_size, batch_size=8, save_results=True,
                               save_viz=False)
        d = root / 'res' / layer
        out[f'{layer}_ids'] = np.loadtxt(d / 'ids.csv', delimiter=',').astype(np.int64)
        out[f'{layer}_activations'] = np.loadtxt(d / 'activations.csv', delimiter=',').astype(np.float32)
        out[f'{layer}_images'] = np.load(d / 'images.npy')
        out[f'{layer}_masks'] = np.load(d / 'masks.npy')
        print(f'exemplars golden [{layer}]: images {out[f"{layer}_images"].shape} masks mean '
              f'{out[f"{layer}_masks"].mean():.4f} ids[0] {out[f"{layer}_ids"][0].tolist()}')
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'exemplars.npz'), **out)


class FeaturesToImage(torch.nn.Module):
    """Toy generator head, as in the reference's own `test_generative` (`tests/exemplars/compute_test.py:265-272`):
    the first three feature channels, squashed to [0, 1], are the "generated image"."""

    def forward(self, features):
        return torch.sigmoid(features[:, :3])


def generative_toy_model(seed: int = 0):
    """The toy CNN with the image head appended: representation in, image out; `conv_2` is the dissected layer."""
    import collections
    layers = list(exemplar_toy_model(seed).named_children())
    layers.append(('output', FeaturesToImage()))
    return torch.nn.Sequential(collections.OrderedDict(layers)).eval()


GENERATIVE_CASES = (('conv_2', 24, 3, None), ('conv_1', 16, 4, (0, 2, 5)))  # (layer, output_size, k, units)


def make_generative_golden():
    """The reference's `exemplars.compute.generative` (`src/exemplars/compute.py:352-437`) on the toy generator: the
    top images are the model's OUTPUTS for the top-activating representations, renormalised from [0, 1] to bytes
    (no dataset normaliser: `renormalize.renormalizer(source=dataset)` falls back to 'pt')."""
    import tempfile
    from torch.utils import data
    compute = ref_import.import_reference_exemplars()
    model, zs = generative_toy_model(), exemplar_toy_images(seed=11)
    out = {}
    for layer, output_size, k, units in GENERATIVE_CASES:
        root = pathlib.Path(tempfile.mkdtemp())
        compute.generative(model, data.TensorDataset(zs), layer, device='cpu', results_dir=root / 'res',
                           viz_dir=root / 'viz', display_progress=False, num_workers=0, k=k, quantile=0.99,
                           image_size=16, output_size=output_size, batch_size=8, save_results=True, save_viz=False,
                           units=units)
        d = root / 'res' / layer
        out[f'{layer}_ids'] = np.loadtxt(d / 'ids.csv', delimiter=',').astype(np.int64)
        out[f'{layer}_activations'] = np.loadtxt(d / 'activations.csv', delimiter=',').astype(np.float32)
        out[f'{layer}_images'] = np.load(d / 'images.npy')
        out[f'{layer}_masks'] = np.load(d / 'masks.npy')
        if units is not None:
            out[f'{layer}_units'] = np.load(d / 'units.npy')
        print(f'generative golden [{layer}]: images {out[f"{layer}_images"].shape} mean {out[f"{layer}_images"].mean():.2f} '
              f'masks mean {out[f"{layer}_masks"].mean():.4f} ids[0] {out[f"{layer}_ids"][0].tolist()}')
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'exemplars_generative.npz'), **out)


class ToyViT(torch.nn.Module):
    """A one-block ViT in the shape of DINO's `vit_small(patch_size=8)` blocks (`blocks.N.mlp.fc1` is what the
    reference dissects, `src/exemplars/models.py:236-247`): patch embedding + CLS token, one pre-norm block whose MLP
    hidden layer `mlp.fc1` yields (batch, 1 + patches, units)."""

    def __init__(self, seed: int = 0, size: int = 16, patch: int = 4, dim: int = 12, units: int = 10):
        super().__init__()
        self.patch_embed = torch.nn.Conv2d(3, dim, patch, stride=patch)
        self.cls_token = torch.nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = torch.nn.Parameter(torch.zeros(1, 1 + (size // patch) ** 2, dim))
        self.norm1 = torch.nn.LayerNorm(dim)
        self.attn = torch.nn.MultiheadAttention(dim, 2, batch_first=True)
        self.norm2 = torch.nn.LayerNorm(dim)
        self.mlp = torch.nn.Sequential(collections_ordered([('fc1', torch.nn.Linear(dim, units)), ('act', torch.nn.GELU()),
                                                            ('fc2', torch.nn.Linear(units, dim))]))
        gen = torch.Generator().manual_seed(seed + 31)
        with torch.no_grad():
            for param in self.parameters():
                param.copy_(torch.randn(param.shape, generator=gen) * 0.4)

    def forward(self, images):
        tokens = self.patch_embed(images).flatten(2).transpose(1, 2)
        tokens = torch.cat([self.cls_token.expand(len(tokens), -1, -1), tokens], dim=1) + self.pos_embed
        normed = self.norm1(tokens)
        tokens = tokens + self.attn(normed, normed, normed, need_weights=False)[0]
        return tokens + self.mlp(self.norm2(tokens))


def collections_ordered(pairs):
    import collections
    return collections.OrderedDict(pairs)


VIT_CASE = ('mlp.fc1', 16, 3)  # (layer, output_size, k)


def make_vit_exemplars_golden():
    """The reference's `discriminative` with `transform_hiddens=spatialize_vit_mlp` (`src/exemplars/transforms.py:
    55-81`, the DINO ViT-S/8 configuration of `src/exemplars/models.py:236-247`) on the toy ViT."""
    import importlib
    import tempfile
    from torch.utils import data
    compute = ref_import.import_reference_exemplars()
    ref_transforms = importlib.import_module('src.exemplars.transforms')
    model, images = ToyViT().eval(), exemplar_toy_images(seed=23)
    layer, output_size, k = VIT_CASE
    root = pathlib.Path(tempfile.mkdtemp())
    compute.discriminative(model, data.TensorDataset(images), layer=layer, device='cpu', results_dir=root / 'res',
                           viz_dir=root / 'viz', display_progress=False, num_workers=0, k=k, quantile=0.99,
                           image_size=16, output_size=output_size, batch_size=8, save_results=True, save_viz=False,
                           transform_hiddens=ref_transforms.spatialize_vit_mlp)
    d = root / 'res' / layer
    out = {'ids': np.loadtxt(d / 'ids.csv', delimiter=',').astype(np.int64),
           'activations': np.loadtxt(d / 'activations.csv', delimiter=',').astype(np.float32),
           'images': np.load(d / 'images.npy'), 'masks': np.load(d / 'masks.npy')}
    print(f'vit exemplars golden: images {out["images"].shape} masks mean {out["masks"].mean():.4f} ids[0] {out["ids"][0].tolist()}')
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'exemplars_vit.npz'), **out)


def reranker_similarity(images, texts, masks=None):
    """Deterministic stand-in for `CLIPWithMasks.forward` (`src/milan/rerankers.py:141-229`): (k, n) scores from the
    exemplars, the candidate texts and, when given, the masks."""
    per_image = images.float().mean(dim=(1, 2, 3))
    lengths = torch.tensor([float(len(text)) for text in texts])
    sim = torch.sin(per_image[:, None] * 7.0 + lengths[None, :] * 1.3)
    if masks is not None:
        sim = sim + 0.5 * torch.cos(masks.float().mean(dim=(1, 2, 3))[:, None] * 5.0 + lengths[None, :] * 0.7)
    return sim


def reranker_inputs():
    gen = torch.Generator().manual_seed(41)
    images = torch.rand(3, 4, 3, 8, 8, generator=gen)
    masks = (torch.rand(3, 4, 1, 8, 8, generator=gen) > 0.7).float()
    words = ('edges', 'of', 'round', 'objects', 'text', 'sky', 'animal', 'faces', 'red', 'and', 'green', 'stripes')
    texts = []
    for n in (5, 7, 2):
        texts.append(tuple(' '.join(words[int(i)] for i in torch.randint(0, len(words), (int(torch.randint(1, 6, (1,), generator=gen)),),
                                                                              generator=gen)) for _ in range(n)))
    return images, masks, tuple(texts)


def make_reranker_golden():
    """The reference's `CLIPWithMasksReranker.forward` (`src/milan/rerankers.py:261-330`) around the stand-in
    similarity: what `milan.rerankers.SimilarityReranker` has to reproduce."""
    import importlib
    import json
    ref_import.import_reference()
    ref_rerankers = importlib.import_module('src.milan.rerankers')
    images, masks, texts = reranker_inputs()
    out = {}
    for name, default_lam, lam in (('default', .5, None), ('lam0.2', .5, .2), ('unmasked_only', 1., None)):
        reranker = ref_rerankers.CLIPWithMasksReranker(reranker_similarity, lam=default_lam)
        got = reranker(images, masks, texts, lam=lam)
        out[name] = {'default_lam': default_lam, 'lam': lam, 'texts': [list(t) for t in got.texts],
                     'orders': [list(o) for o in got.orders], 'scores': [list(s) for s in got.scores]}
    with open(os.path.join(GOLDEN_DIR, 'reranker.json'), 'w') as handle:
        json.dump(out, handle, indent=0, sort_keys=True)
    print('reranker golden:', {k: v['orders'][0] for k, v in out.items()})


def payload_skeleton(value):
    """A checkpoint payload with every tensor replaced by ['tensor', shape, dtype] (JSON-serialisable)."""
    if isinstance(value, dict):
        return {str(key): payload_skeleton(item) for key, item in value.items()}
    if torch.is_tensor(value):
        return ['tensor', list(value.shape), str(value.dtype)]
    if isinstance(value, (tuple, list)):
        return [payload_skeleton(item) for item in value]
    return value


def make_checkpoint_skeleton(milan, lang):
    """Structure of `Decoder.serialize()` (`src/utils/serialize.py:80-118,188-219`, `decoders.py:1072-1109`) as the
    reference writes it for the shipped architecture: what `milan.pretrained()` has to ingest."""
    import json
    vocab = synthetic.synthetic_vocab(40)
    indexer = lang.Indexer(lang.Vocab(tuple(vocab)), tokenize=None, start=True, stop=True, pad=True, unk=True)
    decoder = milan.decoders.Decoder(indexer, milan.encoders.PyramidConvEncoder('resnet101', pretrained=False),
                                     lm=milan.lms.LanguageModel(indexer))
    with open(os.path.join(GOLDEN_DIR, 'checkpoint_skeleton.json'), 'w') as handle:
        json.dump(payload_skeleton(decoder.serialize()), handle, indent=0, sort_keys=True)
    print('checkpoint skeleton:', len(decoder.state_dict()), 'state_dict entries')


LANG_TOKENS = ('.', ',', ';', ':', '-', 'dog', 'cat', 'a', 'top', 'x.y', 'end-', 'The', 'of')
LANG_FLAGS = ((True, True, True, True), (False, True, True, True), (True, True, False, False),
              (False, False, False, False))
LANG_UNINDEX_KWARGS = ({}, {'specials': False}, {'start': False, 'unk': False}, {'stop': False, 'pad': False})


def lang_cases(n_specials: int, seed: int):
    """Seeded random id sequences over a vocabulary of LANG_TOKENS + `n_specials` special ids."""
    import random
    rng = random.Random(seed)
    n = len(LANG_TOKENS) + n_specials
    return [[[rng.randrange(n) for _ in range(rng.randint(1, 12))] for _ in range(rng.randint(1, 4))]
            for _ in range(40)]


def make_lang_golden(lang):
    """`Indexer.unindex` / `Indexer.reconstruct` of the unmodified reference (`src/utils/lang.py:573-730`) on
    seeded random sequences, for every combination of enabled specials the positional flag matching cares about."""
    import json
    out = []
    for fi, (start, stop, pad, unk) in enumerate(LANG_FLAGS):
        indexer = lang.Indexer(lang.Vocab(LANG_TOKENS), tokenize=None, start=start, stop=stop, pad=pad, unk=unk)
        for batch in lang_cases(len(indexer.specials), seed=fi):
            out.append({
                'flags': fi,
                'ids': batch,
                'unindex': [[list(seq) for seq in indexer.unindex(batch, **kw)] for kw in LANG_UNINDEX_KWARGS],
                'unindex_single': list(indexer.unindex(batch[0])),
                'reconstruct': list(indexer.reconstruct(batch)),
                'reconstruct_single': indexer.reconstruct(batch[0]),
                'reconstruct_tokens': list(indexer.reconstruct(indexer.unindex(batch))),
            })
    with open(os.path.join(GOLDEN_DIR, 'lang_reconstruct.json'), 'w') as handle:
        json.dump(out, handle, separators=(',', ':'))
    print(f'lang golden: {len(out)} batches')


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    milan, lang = ref_import.import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    vocab = synthetic.synthetic_vocab(5000)
    if '--only-score' in sys.argv:
        return make_score_golden(milan, lang, vocab)
    if '--only-encoders' in sys.argv:
        return make_encoder_variant_goldens(milan)
    if '--only-reranker' in sys.argv:
        return make_reranker_golden()
    if '--only-exemplars' in sys.argv:
        make_generative_golden()
        make_vit_exemplars_golden()
        return make_exemplars_golden()
    if '--only-checkpoint' in sys.argv:
        return make_checkpoint_skeleton(milan, lang)
    if '--only-lang' in sys.argv:
        return make_lang_golden(lang)
    make_score_golden(milan, lang, vocab)
    make_encoder_variant_goldens(milan)
    make_exemplars_golden()
    make_checkpoint_skeleton(milan, lang)

    # ---- encoder golden: reference PyramidConvEncoder('resnet101') on seeded exemplars.
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=3.0)
    decoder = build_reference_decoder(milan, lang, sd, vocab)
    images_u8, masks_u8 = synthetic.synthetic_exemplars(ENC_NEURONS, K, seed=0, zero_mask_fraction=0.1)
    masks_u8[0, 0] = 0  # guarantee one all-zero mask
    masks_u8[1, 3, :, 100:102, 50:52] = 0
    scale = torch.tensor(1.0 / 255.0, dtype=torch.float64).to(torch.float32)
    images = images_u8.float().mul(scale)  # TopImagesDataset contract, src/milannotations/datasets.py:191-197
    masks = masks_u8.float()
    with torch.no_grad():
        features = decoder.encode(images, masks)
        out_greedy = decoder(images, masks, strategy='greedy', mi=False)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'encoder_resnet101.npz'),
                        features=features.numpy(), greedy_tokens=out_greedy.tokens.numpy(),
                        greedy_scores=out_greedy.scores.numpy(),
                        meta=np.array([ENC_NEURONS, K, 0], dtype=np.int64))
    print('encoder golden: features', tuple(features.shape), 'abs mean', features.abs().mean().item(),
          'max', features.abs().max().item())

    make_lang_golden(lang)

    # ---- decoder goldens, three logit regimes, from seeded features.
    feats = synthetic_features(DEC_NEURONS, K, seed=0)
    for name, (sharpen, stop_bias) in VARIANTS.items():
        sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
        decoder = build_reference_decoder(milan, lang, sd, vocab, with_encoder=False)
        stop_index = decoder.indexer.stop_index
        with torch.no_grad():
            state = decoder.init_state(feats, lm=False)
            start = torch.full((DEC_NEURONS,), decoder.indexer.start_index, dtype=torch.long)
            first = decoder.step(feats, start, state)
            greedy = decoder(feats, strategy='greedy', mi=False)
            greedy_mi = decoder(feats, strategy='greedy', mi=True)
            beam = decoder(feats, strategy='beam', mi=False, beam_size=50)
            rerank = decoder(feats, strategy='rerank', beam_size=50)
            small_beam = decoder(feats, strategy='rerank', beam_size=7, length=9)
            beam_mi = decoder(feats, strategy='beam', beam_size=10)  # mi defaults to True (decoders.py:385-387)
            inputs_lm = torch.cat([torch.full((DEC_NEURONS * 50, 1), decoder.indexer.start_index, dtype=torch.long),
                                   beam.beam_tokens.view(DEC_NEURONS * 50, -1)], dim=-1)
            lm_scores = decoder.lm(inputs_lm, reduce=True)
        top_v, top_i = first.predictions.topk(8, dim=-1)
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, f'decoder_{name}.npz'),
            init_h=state.h.numpy(), init_c=state.c.numpy(),
            step0_attn=first.attentions.numpy(), step0_h=first.state.h.numpy(), step0_c=first.state.c.numpy(),
            step0_top_values=top_v.numpy(), step0_top_indices=top_i.numpy(),
            step0_logsumexp=first.predictions.logsumexp(-1).numpy(),
            greedy_tokens=greedy.tokens.numpy(), greedy_scores=greedy.scores.numpy(),
            greedy_attn=greedy.attentions.numpy(),
            greedy_chosen_logp=greedy.predictions.gather(2, greedy.tokens.unsqueeze(-1)).squeeze(-1).numpy(),
            greedy_mi_tokens=greedy_mi.tokens.numpy(), greedy_mi_scores=greedy_mi.scores.numpy(),
            beam_tokens=beam.beam_tokens.numpy(), beam_scores=beam.beam_scores.numpy(),
            rerank_tokens=rerank.tokens.numpy(), rerank_scores=rerank.scores.numpy(),
            small_beam_tokens=small_beam.beam_tokens.numpy(), small_beam_scores=small_beam.beam_scores.numpy(),
            small_rerank_tokens=small_beam.tokens.numpy(), small_rerank_scores=small_beam.scores.numpy(),
            lm_scores=lm_scores.numpy(),
            beam_mi_tokens=beam_mi.beam_tokens.numpy(), beam_mi_scores=beam_mi.beam_scores.numpy(),
            captions=np.array(rerank.captions), greedy_captions=np.array(greedy.captions),
            meta=np.array([DEC_NEURONS, K, stop_index], dtype=np.int64))
        print(f'decoder golden [{name}]: beam T={beam.beam_tokens.shape[-1]} rerank[0]={rerank.captions[0]!r} '
              f'greedy score {greedy.scores[0].item():.4f} beam top {beam.beam_scores[0, 0].item():.4f}')


if __name__ == '__main__':
    main()
