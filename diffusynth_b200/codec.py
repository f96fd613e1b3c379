"""Spectral representation <-> waveform, batched on the GPU.

Replaces the per-sample CPU loop of the reference's decode glue
(webUI/natural_language_guided_4/utils.py:229-241: ``decode_stft`` -> ``depad_STFT`` -> ``librosa.istft``) and
encode glue (sound2sound_with_text.py:80-94: ``adjust_audio_length`` -> ``librosa.stft`` -> ``pad_STFT`` ->
``encode_stft``) with one fused kernel sequence each (ds_stft_decode_istft / ds_stft_encode), fp32."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import check


@torch.no_grad()
def spectrogram_to_waveform(spec: torch.Tensor) -> torch.Tensor:
    """[B,3,512,T] fp32 (log1p-magnitude, cos, sin) -> [B, 256*(T-1)] fp32."""
    spec = spec.float().contiguous()
    assert spec.dim() == 4 and spec.shape[1] == 3 and spec.shape[2] == 512, "expected [B,3,512,T]"
    B, T = spec.shape[0], spec.shape[3]
    lib = _lib.load()
    frames = torch.empty((B, T, 1024), dtype=torch.float32, device=spec.device)
    wave = torch.empty((B, lib.ds_istft_length(T)), dtype=torch.float32, device=spec.device)
    check(lib.ds_stft_decode_istft(spec.data_ptr(), frames.data_ptr(), wave.data_ptr(), B, T, ops._stream()), "ds_stft_decode_istft")
    return wave


def adjust_audio_length(audio: torch.Tensor, desired_length: int) -> torch.Tensor:
    """Crop or zero-pad the last axis (tools.py:126-151, without the resampling branch)."""
    L = audio.shape[-1]
    if L >= desired_length:
        return audio[..., :desired_length]
    return torch.nn.functional.pad(audio, (0, desired_length - L))


@torch.no_grad()
def waveform_to_spectrogram(wave: torch.Tensor, time_resolution: int = 256) -> torch.Tensor:
    """[B, L] fp32 -> [B,3,512,time_resolution]: STFT(1024, hop 256, hann, centre zero-pad), DC row dropped,
    zero-padded in time, encoded as (log1p|D|, cos, sin)."""
    wave = wave.float().contiguous()
    B, L = wave.shape
    spec = torch.empty((B, 3, 512, time_resolution), dtype=torch.float32, device=wave.device)
    check(_lib.load().ds_stft_encode(wave.data_ptr(), L, spec.data_ptr(), B, time_resolution, ops._stream()), "ds_stft_encode")
    return spec


@torch.no_grad()
def encodeBatch2GradioOutput_STFT(decoder, latent_vector_batch, resolution=(512, 256), original_STFT_batch=None):
    """Tensor->waveform part of utils.py:194-267.  Returns the reference's 6-tuple; the two image lists
    (dB spectrogram / phase uint8 renderings, utils.py:237-238) are not produced here and come
    back empty, the signals are float32 numpy arrays like the reference's float64 ones."""
    if original_STFT_batch is not None:
        raise NotImplementedError("original_STFT_batch amplitude swap")
    if isinstance(latent_vector_batch, np.ndarray):
        latent_vector_batch = torch.from_numpy(latent_vector_batch).to(next(decoder.parameters()).device)
    rec = decoder(latent_vector_batch)
    wave = spectrogram_to_waveform(rec).cpu().numpy()
    return [], [], [w for w in wave], [], [], []
