"""GPU: the secondary encoder configs on the CUDA engine, written like the reference's own encoder tests
(`tests/milan/encoders_test.py:25-77`: shape / no-NaN / exact zeros for all-zero masks, for `resnet18` pyramids)
plus numeric parity with goldens produced by the UNMODIFIED reference (`oracle/make_golden.py`)."""
import os

import numpy as np
import pytest
import torch

from neuron_descriptions_b200 import synthetic
from oracle import milan_oracle as O
from oracle.make_golden import encoder_variant_inputs

pytestmark = pytest.mark.gpu

BATCH_SIZE = 10
IMAGE_SIZE = 224
IMAGE_SHAPE = (3, IMAGE_SIZE, IMAGE_SIZE)
MASK_SHAPE = (1, IMAGE_SIZE, IMAGE_SIZE)


def _encoder(kind, config, seed=3):
    from neuron_descriptions_b200.milan import encoders
    cls = encoders.SpatialConvEncoder if kind == 'spatial' else encoders.PyramidConvEncoder
    encoder = cls(config=config, pretrained=False)
    encoder.load_state_dict(synthetic.synthetic_encoder_state_dict(config, seed=seed))
    return encoder.to('cuda:0')


@pytest.fixture
def images():
    return torch.rand(BATCH_SIZE, *IMAGE_SHAPE, generator=torch.Generator().manual_seed(1))


@pytest.fixture
def masks():
    return torch.randint(2, size=(BATCH_SIZE, *MASK_SHAPE), generator=torch.Generator().manual_seed(2)).float()


def test_pyramid_conv_encoder_init_bad_config():
    from neuron_descriptions_b200.milan import encoders
    bad = 'bad-config'
    with pytest.raises(ValueError, match=f'.*{bad}.*'):
        encoders.PyramidConvEncoder(config=bad)
    with pytest.raises(ValueError, match=f'.*{bad}.*'):
        encoders.SpatialConvEncoder(config=bad)


@pytest.mark.parametrize('config', ('resnet18', 'alexnet', 'resnet50'))
def test_pyramid_conv_encoder_forward(config, images, masks):
    encoder = _encoder('pyramid', config)
    actual = encoder(images, masks)
    assert actual.shape == (BATCH_SIZE, *encoder.feature_shape)
    assert not torch.isnan(actual).any()


@pytest.mark.parametrize('config', ('resnet18', 'alexnet', 'resnet50'))
def test_pyramid_conv_encoder_forward_invalid_mask(config, images, masks):
    encoder = _encoder('pyramid', config)
    masks[-2:] = 0
    actual = encoder(images, masks)
    assert actual.shape == (BATCH_SIZE, *encoder.feature_shape)
    assert actual[-2:].eq(0).all()
    assert not actual[:-2].eq(0).all()
    assert not torch.isnan(actual).any()


@pytest.mark.parametrize('config', ('resnet18', 'alexnet'))
def test_pyramid_conv_encoder_forward_all_invalid_masks(config, images, masks):
    encoder = _encoder('pyramid', config)
    actual = encoder(images, torch.zeros_like(masks))
    assert actual.shape == (BATCH_SIZE, *encoder.feature_shape)
    assert actual.eq(0).all()


@pytest.mark.parametrize('kind,config', [('pyramid', 'resnet18'), ('pyramid', 'resnet50'), ('spatial', 'resnet18'),
                                       ('pyramid', 'alexnet')])
def test_encoder_variants_match_reference_golden(golden_dir, kind, config):
    g = np.load(os.path.join(golden_dir, 'encoder_variants.npz'))
    images_u8, masks_u8 = encoder_variant_inputs()
    encoder = _encoder(kind, config)
    feats = encoder(images_u8, masks_u8).cpu()
    ref = torch.from_numpy(g[f'{kind}_{config}'])
    assert feats.shape == ref.shape
    scale = ref.abs().max().item()
    err = (feats - ref).abs().max().item()
    print(f'{kind}/{config}: max abs err {err:.3e} (scale {scale:.3f})')
    assert err <= 1e-3 * scale
    # float inputs take the same path bit for bit
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    assert torch.equal(encoder(images_f, masks_f).cpu(), feats)
    if kind == 'spatial':  # masks default to all ones (src/milan/encoders.py:206-207)
        no_mask = encoder(images_u8, None).cpu()
        ones = encoder(images_u8, torch.ones_like(masks_u8)).cpu()
        assert torch.equal(no_mask, ones)


def test_spatial_decoder_end_to_end():
    """A `Decoder` over `SpatialConvEncoder('resnet18')`: 3 exemplars x 49 positions = 147 keys of size 512
    (`src/milan/decoders.py:546`), greedy + rerank vs the oracle."""
    from neuron_descriptions_b200 import milan
    from neuron_descriptions_b200.milan import lang
    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=2, sharpen=12.0, stop_bias=1.0, feature_size=512, encoder_arch='resnet18')
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.SpatialConvEncoder('resnet18', pretrained=False),
                            lm=milan.LanguageModel(indexer), max_neurons=4)
    decoder.load_state_dict(sd)
    decoder.to('cuda:0')
    images_u8, masks_u8 = synthetic.synthetic_exemplars(2, 3, seed=9)
    images_f, masks_f = O.to_float_inputs(images_u8, masks_u8)
    with torch.no_grad():
        ref_feats = O.encode(images_f, masks_f, sd, arch='resnet18', kind='spatial')
        ref_greedy = O.decode(ref_feats, sd, vocab, strategy='greedy', mi=False)
        ref = O.decode(ref_feats, sd, vocab, strategy='rerank', beam_size=8)
    assert ref_feats.shape == (2, 147, 512)
    feats = decoder.encode(images_f, masks_f)
    assert feats.shape == (2, 147, 512)
    scale = ref_feats.abs().max().item()
    assert (feats.cpu() - ref_feats).abs().max().item() <= 1e-3 * scale
    out = decoder(images_f, masks_f, strategy='greedy', mi=False)
    torch.testing.assert_close(out.scores.cpu(), ref_greedy.scores, atol=1e-3, rtol=0)
    assert torch.equal(out.tokens.cpu(), ref_greedy.tokens)
    assert out.attentions.shape == (2, 15, 147)
    out = decoder(images_f, masks_f, strategy='rerank', beam_size=8)
    torch.testing.assert_close(out.scores.cpu(), ref.scores, atol=1e-3, rtol=0)
    assert torch.equal(out.tokens.cpu(), ref.tokens)
    assert tuple(out.captions) == tuple(ref.captions)
