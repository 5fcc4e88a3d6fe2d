"""Print the measured parity errors (CUDA path vs reference goldens / oracle) — run on the GPU box.

    python scripts/parity_report.py > gpurun_out/parity_report.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neuron_descriptions_b200 import synthetic  # noqa: E402
from neuron_descriptions_b200.engine import Engine  # noqa: E402
from oracle import milan_oracle as O  # noqa: E402
from oracle.make_golden import VARIANTS, synthetic_features  # noqa: E402


def main():
    vocab = synthetic.synthetic_vocab(5000)
    V = len(vocab) + 4
    golden = os.path.join(ROOT, 'tests', 'golden')
    print('quantity, variant, max |cuda - reference|, reference scale, tokens identical')
    for precision in ('split', 'fast'):
        sd = synthetic.synthetic_state_dict(seed=0, sharpen=3.0)
        engine = Engine(sd, vocab_size=V, device='cuda:0', max_neurons=16, precision=precision)
        g = np.load(os.path.join(golden, 'encoder_resnet101.npz'))
        images_u8, masks_u8 = synthetic.synthetic_exemplars(2, 15, seed=0, zero_mask_fraction=0.1)
        masks_u8[0, 0] = 0
        masks_u8[1, 3, :, 100:102, 50:52] = 0
        feats = engine.encode(images_u8.view(-1, 3, 224, 224), masks_u8.view(-1, 1, 224, 224)).cpu().view(2, 15, -1)
        ref = torch.from_numpy(g['features'])
        print(f'[{precision}] encoder features, -, {(feats - ref).abs().max().item():.3e}, {ref.abs().max().item():.3f}, -')
        tok, sc, _, _ = engine.decode_greedy(feats.cuda(), 15, mi=False, temperature=0.2)
        print(f'[{precision}] encoder->greedy (e2e), -, {np.abs(sc.cpu().numpy() - g["greedy_scores"]).max():.3e}, '
              f'{np.abs(g["greedy_scores"]).max():.1f}, {np.array_equal(tok.cpu().numpy(), g["greedy_tokens"])}')
        engine.close()
        for name, (sharpen, stop_bias) in sorted(VARIANTS.items()):
            sd = synthetic.synthetic_state_dict(seed=0, sharpen=sharpen, stop_bias=stop_bias, with_encoder=False)
            engine = Engine(sd, vocab_size=V, device='cuda:0', max_neurons=16, precision=precision)
            g = np.load(os.path.join(golden, f'decoder_{name}.npz'))
            n, k, stop = g['meta'].tolist()
            feats = synthetic_features(n, k, seed=0)
            tok, sc, pred, attn = engine.decode_greedy(feats, 15, mi=False, temperature=0.2)
            print(f'[{precision}] greedy score, {name}, {np.abs(sc.cpu().numpy() - g["greedy_scores"]).max():.3e}, '
                  f'{np.abs(g["greedy_scores"]).max():.1f}, {np.array_equal(tok.cpu().numpy(), g["greedy_tokens"])}')
            tok, sc, _, _ = engine.decode_greedy(feats, 15, mi=True, temperature=0.2)
            print(f'[{precision}] greedy-MI score, {name}, {np.abs(sc.cpu().numpy() - g["greedy_mi_scores"]).max():.3e}, '
                  f'{np.abs(g["greedy_mi_scores"]).max():.1f}, {np.array_equal(tok.cpu().numpy(), g["greedy_mi_tokens"])}')
            bt, bs, steps, tok, sc, lm = engine.decode_beam(feats, 15, 50, True, 0.2)
            T = int(steps[0])
            same = T == g['beam_tokens'].shape[-1] and np.array_equal(bt[..., :T].cpu().numpy(), g['beam_tokens'])
            print(f'[{precision}] beam scores (50), {name}, {np.abs(bs.cpu().numpy() - g["beam_scores"]).max():.3e}, '
                  f'{np.abs(g["beam_scores"]).max():.1f}, {same}')
            print(f'[{precision}] LM scores, {name}, {np.abs(lm.view(-1).cpu().numpy() - g["lm_scores"]).max():.3e}, '
                  f'{np.abs(g["lm_scores"]).max():.1f}, -')
            same = T == g['rerank_tokens'].shape[-1] and np.array_equal(tok[..., :T].cpu().numpy(), g['rerank_tokens'])
            print(f'[{precision}] rerank score, {name}, {np.abs(sc.cpu().numpy() - g["rerank_scores"]).max():.3e}, '
                  f'{np.abs(g["rerank_scores"]).max():.1f}, {same}')
            engine.close()


if __name__ == '__main__':
    main()
