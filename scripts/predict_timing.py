"""Time the user-facing `Decoder.predict(dataset)` (mmapped uint8 exemplar files -> captions) against the raw
`milan_describe_host` C-ABI call on the same neurons (GPU box). Measurement aid, not a bench number."""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_descriptions_b200 import milan, milannotations, synthetic  # noqa: E402
from neuron_descriptions_b200.milan import lang  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vocab = synthetic.synthetic_vocab(5000)
sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0)
root = tempfile.mkdtemp(prefix='milan_exemplars_')
os.makedirs(os.path.join(root, 'layer4'))
images, masks = synthetic.synthetic_exemplars(N, 15, seed=1)
np.save(os.path.join(root, 'layer4', 'images.npy'), images.numpy())
np.save(os.path.join(root, 'layer4', 'masks.npy'), masks.numpy())
dataset = milannotations.TopImagesDataset(root)
indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                        lm=milan.LanguageModel(indexer), max_neurons=64)
decoder.load_state_dict(sd)
decoder.to('cuda:0')
for label in ('warm-up', 'timed'):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    captions = decoder.predict(dataset, strategy='rerank', temperature=.2, beam_size=50, device='cuda:0',
                               display_progress_as=None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'predict [{label}]: {N} neurons in {dt:.3f} s = {N / dt:.1f} neurons/s; first caption {captions[0]!r}')
pinned = (images.pin_memory(), masks.pin_memory())
for label in ('warm-up', 'timed'):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tokens, scores, steps = decoder.engine.describe_host(*pinned, strategy='rerank')
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'describe_host [{label}]: {N} neurons in {dt:.3f} s = {N / dt:.1f} neurons/s')
ref = decoder.last_predict_tokens
print('predict tokens == describe_host tokens:', bool(torch.equal(ref, tokens)))
