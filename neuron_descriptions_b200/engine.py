"""Torch-tensor front end of the C ABI: owns one `MilanEngine*` on one GPU.

PyTorch is only plumbing here (device memory + current stream); all compute happens in libmilan_b200.so.
"""
import ctypes
from typing import Mapping, Optional, Tuple

import torch

from neuron_descriptions_b200 import _lib


def _ptr(tensor: Optional[torch.Tensor]):
    return None if tensor is None else ctypes.c_void_p(tensor.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One engine per (process, GPU): weights + workspace live on `device`."""

    def __init__(self,
                 state_dict: Mapping[str, torch.Tensor],
                 vocab_size: int,
                 device,
                 embedding_size: int = 128,
                 hidden_size: int = 512,
                 attention_size: Optional[int] = None,
                 feature_size: int = 3904,
                 lm_embedding_size: int = 128,
                 lm_hidden_size: int = 512,
                 precision: str = 'split',
                 max_neurons: int = 32,
                 max_beam: int = 50,
                 max_keys: int = 15,
                 max_length: int = 15,
                 encoder_arch: str = 'resnet101',
                 encoder_kind: str = 'pyramid',
                 max_images: Optional[int] = None,
                 decoder: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError('milan_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError(f'milan_b200 engine cannot run on device {device!r}; it is CUDA-only')
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', index)
        has_encoder = any(key.startswith('encoder.encoder.model.') for key in state_dict)
        has_lm = any(key.startswith('lm.') for key in state_dict)
        if encoder_arch not in _lib.ENCODER_ARCHS or encoder_kind not in _lib.ENCODER_KINDS:
            raise ValueError(f'encoder not supported: {encoder_kind}/{encoder_arch}')
        self.keys_per_image = 49 if encoder_kind == 'spatial' else 1
        if max_images is None:
            max_images = max(1, max_neurons * max_keys // self.keys_per_image)
        cfg = _lib.MilanConfig(
            vocab_size=vocab_size, embedding_size=embedding_size, hidden_size=hidden_size,
            attention_size=attention_size or min(hidden_size, feature_size), feature_size=feature_size,
            start_index=vocab_size - 4, stop_index=vocab_size - 3, has_encoder=int(has_encoder), has_lm=int(has_lm),
            lm_embedding_size=lm_embedding_size, lm_hidden_size=lm_hidden_size,
            precision={'split': _lib.PRECISION_SPLIT, 'fast': _lib.PRECISION_FAST}[precision],
            max_images=max_images, max_neurons=max_neurons, max_beam=max_beam, max_keys=max_keys,
            max_length=max_length, encoder_arch=_lib.ENCODER_ARCHS[encoder_arch],
            encoder_kind=_lib.ENCODER_KINDS[encoder_kind])
        self.cfg = cfg
        self.precision = precision
        self.has_encoder, self.has_lm = has_encoder, has_lm
        handle = ctypes.c_void_p()
        _lib.check(self.lib.milan_engine_create(ctypes.byref(cfg), index, ctypes.byref(handle)))
        self.handle = handle
        for name, tensor in state_dict.items():
            if not torch.is_floating_point(tensor):
                continue  # num_batches_tracked
            if not decoder and not name.startswith('encoder.'):
                continue  # encoder-only engine
            host = tensor.detach().to('cpu', torch.float32).contiguous()
            shape = (ctypes.c_int64 * max(host.dim(), 1))(*host.shape)
            _lib.check(self.lib.milan_engine_set_tensor(handle, name.encode(), ctypes.c_void_p(host.data_ptr()), shape,
                                                        host.dim()))
        _lib.check(self.lib.milan_engine_finalize(handle))

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.milan_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max_rows(self) -> int:
        """Decoder / LM rows one call may carry (the engine's workspace holds max_neurons x max_beam beam rows)."""
        return int(self.cfg.max_neurons) * int(self.cfg.max_beam)

    # ------------------------------------------------------------------ helpers
    def _f32(self, tensor: torch.Tensor) -> torch.Tensor:
        return tensor.to(self.device, torch.float32).contiguous()

    def _new(self, *shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ C ABI wrappers
    def encode(self, images: torch.Tensor, masks: Optional[torch.Tensor]) -> torch.Tensor:
        """images (n,3,224,224) uint8|float, masks (n,1,224,224) same dtype family -> (n, F) for a pyramid
        encoder, (n, 49, F) for a spatial one."""
        if images.shape[1:] != (3, 224, 224):
            raise ValueError(f'milan_b200 encoder expects (n,3,224,224) images, got {tuple(images.shape)}')
        if images.dtype == torch.uint8:
            dtype = _lib.DTYPE_U8
            images = images.to(self.device).contiguous()
            if masks is not None:
                masks = masks.to(self.device, torch.uint8).contiguous()
        else:
            dtype = _lib.DTYPE_F32
            images = self._f32(images)
            if masks is not None:
                masks = self._f32(masks)
        n = images.shape[0]
        if self.keys_per_image > 1:
            out = self._new(n, self.keys_per_image, self.cfg.feature_size)
        else:
            out = self._new(n, self.cfg.feature_size)
        _lib.check(self.lib.milan_encode(self.handle, _ptr(images), _ptr(masks), n, dtype, _ptr(out),
                                         _stream(self.device)))
        return out

    def init_state(self, features: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        features = self._f32(features)
        B, n_keys, _ = features.shape
        h, c = self._new(B, self.cfg.hidden_size), self._new(B, self.cfg.hidden_size)
        _lib.check(self.lib.milan_init_state(self.handle, _ptr(features), B, n_keys, _ptr(h), _ptr(c),
                                             _stream(self.device)))
        return h, c

    def step(self, features, tokens, h, c, h_lm=None, c_lm=None, temperature=0.2):
        features = self._f32(features)
        R, n_keys, _ = features.shape
        tokens = tokens.to(self.device, torch.long).contiguous()
        h, c = self._f32(h).clone(), self._f32(c).clone()
        if h_lm is not None:
            h_lm, c_lm = self._f32(h_lm).clone(), self._f32(c_lm).clone()
        predictions = self._new(R, self.cfg.vocab_size)
        attentions = self._new(R, n_keys)
        _lib.check(self.lib.milan_step(self.handle, _ptr(features), n_keys, _ptr(tokens), _ptr(h), _ptr(c), _ptr(h_lm),
                                       _ptr(c_lm), R, 1, float(temperature), _ptr(predictions), _ptr(attentions),
                                       _stream(self.device)))
        return predictions, attentions, h, c, h_lm, c_lm

    def decode_greedy(self, features, length, mi, temperature, forced=None, want_predictions=True):
        features = self._f32(features)
        B, n_keys, _ = features.shape
        tokens = self._new(B, length, dtype=torch.long)
        scores = self._new(B)
        predictions = self._new(B, length, self.cfg.vocab_size) if want_predictions else None
        attentions = self._new(B, length, n_keys)
        if forced is not None:
            forced = forced.to(self.device, torch.long).contiguous()
        _lib.check(self.lib.milan_decode_greedy(self.handle, _ptr(features), B, n_keys, length, int(bool(mi)),
                                                float(temperature), _ptr(forced), _ptr(tokens), _ptr(scores),
                                                _ptr(predictions), _ptr(attentions), _stream(self.device)))
        return tokens, scores, predictions, attentions

    def decode_beam(self, features, length, beam, rerank, temperature, group_size=None, mi=False):
        features = self._f32(features)
        B, n_keys, _ = features.shape
        group_size = group_size or B
        groups = (B + group_size - 1) // group_size
        beam_tokens = self._new(B, beam, length, dtype=torch.long)
        beam_scores = self._new(B, beam)
        steps = self._new(groups, dtype=torch.int32)
        tokens = self._new(B, length, dtype=torch.long)
        scores = self._new(B)
        lm_scores = self._new(B, beam) if rerank else None
        _lib.check(self.lib.milan_decode_beam(self.handle, _ptr(features), B, n_keys, length, beam, group_size,
                                              int(bool(rerank)), int(bool(mi)), float(temperature), _ptr(beam_tokens),
                                              _ptr(beam_scores), _ptr(steps), _ptr(tokens), _ptr(scores),
                                              _ptr(lm_scores), _stream(self.device)))
        return beam_tokens, beam_scores, steps, tokens, scores, lm_scores

    def lm_score(self, inputs: torch.Tensor) -> torch.Tensor:
        inputs = inputs.to(self.device, torch.long).contiguous()
        M, T1 = inputs.shape
        out = self._new(M)
        _lib.check(self.lib.milan_lm_score(self.handle, _ptr(inputs), M, T1, _ptr(out), _stream(self.device)))
        return out

    def lm_logprobs(self, inputs: torch.Tensor) -> torch.Tensor:
        """(M, T) token ids -> (M, T, V) log-probabilities of the next token (`LanguageModel.forward(reduce=False)`)."""
        inputs = inputs.to(self.device, torch.long).contiguous()
        M, T = inputs.shape
        out = self._new(M, T, self.cfg.vocab_size)
        _lib.check(self.lib.milan_lm_logprobs(self.handle, _ptr(inputs), M, T, _ptr(out), _stream(self.device)))
        return out

    def describe_host(self, images_u8: torch.Tensor, masks_u8: torch.Tensor, strategy: str = 'rerank', mi=False,
                      length: int = 15, beam: int = 50, group_size: int = 16, temperature: float = 0.2):
        """HOST uint8 tensors (n,k,3,224,224)/(n,k,1,224,224) -> host (tokens, scores, steps). H2D/D2H inside."""
        assert images_u8.dtype == torch.uint8 and masks_u8.dtype == torch.uint8
        assert images_u8.device.type == 'cpu' and masks_u8.device.type == 'cpu'
        images_u8, masks_u8 = images_u8.contiguous(), masks_u8.contiguous()
        n, k = images_u8.shape[:2]
        tokens = torch.empty(n, length, dtype=torch.long)
        scores = torch.empty(n, dtype=torch.float32)
        steps = torch.empty(n, dtype=torch.int32)
        code = {'greedy': _lib.STRATEGY_GREEDY, 'beam': _lib.STRATEGY_BEAM, 'rerank': _lib.STRATEGY_RERANK}[strategy]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.milan_describe_host(self.handle, _ptr(images_u8), _ptr(masks_u8), n, k, code,
                                                    int(bool(mi)), length, beam, group_size, float(temperature),
                                                    _ptr(tokens), _ptr(scores), _ptr(steps), _stream(self.device)))
        return tokens, scores, steps

    def describe_device(self, images_u8: torch.Tensor, masks_u8: torch.Tensor, strategy: str = 'rerank', mi=False,
                        length: int = 15, beam: int = 50, group_size: int = 16, temperature: float = 0.2):
        """DEVICE uint8 tensors (n,k,3,224,224)/(n,k,1,224,224) -> device (tokens (n,length), scores (n)); the
        chunks are pipelined like `describe_host` (decode of chunk i under the encoder of chunk i+1)."""
        assert images_u8.dtype == torch.uint8 and masks_u8.dtype == torch.uint8
        images_u8 = images_u8.to(self.device).contiguous()
        masks_u8 = masks_u8.to(self.device).contiguous()
        n, k = images_u8.shape[:2]
        tokens = self._new(n, length, dtype=torch.long)
        scores = self._new(n)
        code = {'greedy': _lib.STRATEGY_GREEDY, 'beam': _lib.STRATEGY_BEAM, 'rerank': _lib.STRATEGY_RERANK}[strategy]
        _lib.check(self.lib.milan_describe_device(self.handle, _ptr(images_u8), _ptr(masks_u8), n, k, code,
                                                  int(bool(mi)), length, beam, group_size, float(temperature),
                                                  _ptr(tokens), _ptr(scores), _stream(self.device)))
        return tokens, scores

    def set_profiling(self, enabled: bool):
        _lib.check(self.lib.milan_set_profiling(self.handle, int(enabled)))

    def get_profile(self):
        conv, enc, dec = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        launches = ctypes.c_int64()
        _lib.check(self.lib.milan_get_profile(self.handle, ctypes.byref(conv), ctypes.byref(enc), ctypes.byref(dec),
                                              ctypes.byref(launches)))
        return {'conv_ms': conv.value, 'encoder_ms': enc.value, 'decoder_ms': dec.value,
                'conv_launches': launches.value}
