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


def adjust_audio_length(audio, desired_length: int, original_sample_rate: Optional[int] = None, target_sample_rate: Optional[int] = None):
    """tools.adjust_audio_length (tools.py:126-151): resample when the rates differ, then crop or zero-pad the last axis.
    Tensors stay tensors (the batched path of TextToTimbre.encode_audio), numpy arrays stay numpy arrays (drop-in use).
    The reference resamples with ``librosa.core.resample`` (absent from this image, un-pinned in requirements.txt:3; its default
    ``res_type`` is a band-limited sinc method): here a polyphase FIR resampler (``scipy.signal.resample_poly``) on the host
    -- parity of this branch is UNPINNED by necessity; it is one clip of host-side preprocessing, not part of the device path."""
    is_np = isinstance(audio, np.ndarray)
    if original_sample_rate is not None and target_sample_rate is not None and original_sample_rate != target_sample_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(original_sample_rate), int(target_sample_rate))
        a = audio if is_np else audio.detach().cpu().numpy()
        a = resample_poly(a, int(target_sample_rate) // g, int(original_sample_rate) // g, axis=-1)
        audio = a if is_np else torch.from_numpy(np.ascontiguousarray(a)).to(audio.device, audio.dtype)
    L = audio.shape[-1]
    if L >= desired_length:
        return audio[..., :desired_length]
    if is_np:
        out = np.zeros(audio.shape[:-1] + (desired_length,))          # the reference pads into np.zeros (float64), :146-148
        out[..., :L] = audio
        return out
    return torch.nn.functional.pad(audio, (0, desired_length - L))


# ---- numpy drop-ins of the reference's STFT+ helpers (tools.py:170-191, 320-345) ------------------------------------------------------
def pad_STFT(D: np.ndarray, time_resolution: Optional[int] = 256) -> np.ndarray:
    """tools.pad_STFT (:170-182): drop the DC row, zero-pad (never crop) the time axis to ``time_resolution``."""
    D = D[1:, :]
    if time_resolution is None:
        return D
    pad = time_resolution - D.shape[1]
    return np.pad(D, ((0, 0), (0, pad)), "constant") if pad > 0 else D


def depad_STFT(D_padded: np.ndarray) -> np.ndarray:
    """tools.depad_STFT (:185-191): prepend a zero DC row (float64 zeros, so complex64 input is promoted like the reference's)."""
    return np.concatenate([np.zeros((1, D_padded.shape[1])), D_padded], axis=0)


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def decode_stft(encoded_D: np.ndarray) -> np.ndarray:
    """tools.decode_stft (:334-345): [3, F, T] (log1p magnitude, cos, sin) -> complex [F, T], evaluated by ds_decode_stft in the
    precision of the input (float32 -> complex64, float64 -> complex128), like numpy."""
    enc = np.ascontiguousarray(encoded_D)
    dbl = enc.dtype == np.float64
    if not dbl:
        enc = enc.astype(np.float32, copy=False)
    t = torch.from_numpy(enc).to(_device())
    plane = int(np.prod(enc.shape[1:]))
    out = torch.empty(tuple(enc.shape[1:]) + (2,), dtype=t.dtype, device=t.device)
    check(_lib.load().ds_decode_stft(t.data_ptr(), out.data_ptr(), plane, 1 if dbl else 0, ops._stream()), "ds_decode_stft")
    return torch.view_as_complex(out).cpu().numpy()


def encode_stft(D: np.ndarray) -> np.ndarray:
    """tools.encode_stft (:320-331): complex [F, T] -> [3, F, T] (log1p|D|, cos angle, sin angle) by ds_encode_stft."""
    Dc = np.ascontiguousarray(D)
    dbl = Dc.dtype == np.complex128
    if not dbl:
        Dc = Dc.astype(np.complex64, copy=False)
    t = torch.view_as_real(torch.from_numpy(Dc)).contiguous().to(_device())
    plane = int(np.prod(Dc.shape))
    out = torch.empty((3,) + tuple(Dc.shape), dtype=t.dtype, device=t.device)
    check(_lib.load().ds_encode_stft(t.data_ptr(), out.data_ptr(), plane, 1 if dbl else 0, ops._stream()), "ds_encode_stft")
    return out.cpu().numpy()


def istft(D: np.ndarray, hop_length: int = 256, win_length: int = 1024) -> np.ndarray:
    """``librosa.istft(D, hop_length=256, win_length=1024)`` as the reference calls it (utils.py:184,241,260) for one complex
    [513, T] matrix (DC row ignored: depad_STFT made it zero), on the device through the fused decode + iSTFT kernel."""
    assert hop_length == 256 and win_length == 1024 and D.shape[0] == 513, "the path's iSTFT is fixed at n_fft 1024 / hop 256"
    enc = encode_stft(np.asarray(D)[1:].astype(np.complex64))
    return spectrogram_to_waveform(torch.from_numpy(enc)[None].to(_device()))[0].cpu().numpy()


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
def griffinlim(magnitude: torch.Tensor, n_iter: int = 32, momentum: float = 0.99, init_phase: Optional[torch.Tensor] = None,
               generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Magnitude-only reconstruction: ``librosa.griffinlim(abs_spec, n_iter=n_iter, hop_length=256, win_length=1024)`` as the
    reference's helpers call it (tools.py:63-76 ``save_results``, :194-217 ``nnData2Audio``, :220-223 ``amp_to_audio``;
    librosa defaults: momentum 0.99, random initial phases, centre zero padding), batched on the GPU from the fused STFT kernels.
    ``magnitude`` [B,512,T] holds bins 1..512 of the [513,T] matrix the reference builds (its DC row is zero, :205-210).
    ``init_phase`` [B,512,T] in radians replaces the random start (librosa draws it from an unseeded numpy generator)."""
    mag = magnitude.float().contiguous()
    assert mag.dim() == 3 and mag.shape[1] == 512, "expected [B,512,T]"
    B, _, T = mag.shape
    dev = mag.device
    if init_phase is None:
        init_phase = 2 * np.pi * torch.rand((B, 512, T), device=dev, generator=generator)
    ph = init_phase.to(dev, torch.float32)
    spec = torch.stack([torch.log1p(mag), torch.cos(ph), torch.sin(ph)], dim=1).contiguous()
    rebuilt = torch.empty_like(spec)
    tprev = torch.empty((B, 512, T, 2), dtype=torch.float32, device=dev)
    frames = torch.empty((B, T, 1024), dtype=torch.float32, device=dev)
    lib = _lib.load()
    L = lib.ds_istft_length(T)
    wave = torch.empty((B, L), dtype=torch.float32, device=dev)
    st = ops._stream()
    for it in range(n_iter):
        check(lib.ds_stft_decode_istft(spec.data_ptr(), frames.data_ptr(), wave.data_ptr(), B, T, st), "ds_stft_decode_istft")
        check(lib.ds_stft_encode(wave.data_ptr(), L, rebuilt.data_ptr(), B, T, st), "ds_stft_encode")
        check(lib.ds_griffinlim_update(rebuilt.data_ptr(), tprev.data_ptr(), spec.data_ptr(), float(momentum), 1 if it == 0 else 0, B, T, st),
              "ds_griffinlim_update")
    check(lib.ds_stft_decode_istft(spec.data_ptr(), frames.data_ptr(), wave.data_ptr(), B, T, st), "ds_stft_decode_istft")
    return wave


@torch.no_grad()
def spectrogram_images(spec: torch.Tensor):
    """[B,3,512,T] fp32 spectral representation -> (dB-magnitude image, phase image), uint8 [B,513,T,3] each: what
    ``spectrogram_to_Gradio_image(np.abs(D))`` / ``phase_to_Gradio_image(np.angle(D))`` (utils.py:8-91) give for
    ``D = depad_STFT(decode_stft(spec[b]))`` (utils.py:229-238), for the whole batch in two kernels."""
    spec = spec.float().contiguous()
    assert spec.dim() == 4 and spec.shape[1] == 3 and spec.shape[2] == 512, "expected [B,3,512,T]"
    B, T = spec.shape[0], spec.shape[3]
    mag = torch.empty((B, 513, T, 3), dtype=torch.uint8, device=spec.device)
    ph = torch.empty_like(mag)
    scratch = torch.empty((B,), dtype=torch.int64, device=spec.device)
    check(_lib.load().ds_spec_images(spec.data_ptr(), mag.data_ptr(), ph.data_ptr(), scratch.data_ptr(), B, T, ops._stream()),
          "ds_spec_images")
    return mag, ph


@torch.no_grad()
def latent_images(latents: torch.Tensor) -> torch.Tensor:
    """[B,4,H,W] fp32 -> uint8 [B,8H,8W,4] (latent_representation_to_Gradio_image, utils.py:94-128, batched;
    the input is left untouched -- the reference normalises its numpy argument in place)."""
    lat = latents.float().contiguous()
    assert lat.dim() == 4 and lat.shape[1] == 4, "expected [B,4,H,W]"
    B, _, H, W = lat.shape
    img = torch.empty((B, 8 * H, 8 * W, 4), dtype=torch.uint8, device=lat.device)
    scratch = torch.empty((B, 4, 2), dtype=torch.float32, device=lat.device)
    check(_lib.load().ds_latent_image(lat.data_ptr(), img.data_ptr(), scratch.data_ptr(), B, H, W, ops._stream()), "ds_latent_image")
    return img


def latent_representation_to_Gradio_image(latent_representation) -> np.ndarray:
    """Drop-in for utils.py:94-128: one latent [4,H,W] (tensor or numpy) -> uint8 [8H,8W,4] numpy."""
    if isinstance(latent_representation, np.ndarray):
        latent_representation = torch.from_numpy(latent_representation)
    dev = latent_representation.device if latent_representation.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return latent_images(latent_representation.to(dev)[None])[0].cpu().numpy()


def _decode_lists(spec: torch.Tensor):
    mag, ph = spectrogram_images(spec)
    wave = spectrogram_to_waveform(spec)
    mag, ph, wave = mag.cpu().numpy(), ph.cpu().numpy(), wave.cpu().numpy()
    return [m for m in mag], [p for p in ph], [w for w in wave]


@torch.no_grad()
def encodeBatch2GradioOutput_STFT(decoder, latent_vector_batch, resolution=(512, 256), original_STFT_batch=None):
    """utils.py:194-267 without the per-sample CPU loop.  Returns the reference's 6 lists
    (dB-spectrogram images, phase images, signals, and the same three with channel 0 of the reconstruction replaced by
    ``original_STFT_batch``'s when that is given, else empty).  Images are uint8 [513,T,3] numpy arrays; signals are
    float32 numpy arrays (the reference's are float64)."""
    if isinstance(latent_vector_batch, np.ndarray):
        latent_vector_batch = torch.from_numpy(latent_vector_batch).to(next(decoder.parameters()).device)
    rec = decoder(latent_vector_batch)
    imgs, phases, signals = _decode_lists(rec)
    if original_STFT_batch is None:
        return imgs, phases, signals, [], [], []
    if isinstance(original_STFT_batch, np.ndarray):
        original_STFT_batch = torch.from_numpy(original_STFT_batch)
    swapped = rec.float().clone()
    swapped[:, 0] = original_STFT_batch[:, 0].to(rec.device, torch.float32)          # utils.py:251
    return (imgs, phases, signals) + _decode_lists(swapped)


@torch.no_grad()
def InputBatch2Encode_STFT(encoder, STFT_batch, resolution=(512, 256), quantizer=None, squared=True):
    """utils.py:131-191: encode a batch of spectral representations (and quantise), and render the INPUT batch back to
    images / signals.  ``quantizer=None`` is the reference's VAE branch (:162-164): the encoder must then return
    ``(mu, logvar, z)`` and there is no quantised batch; the deployed VQGAN encoder returns a single tensor, so -- exactly as
    in the reference -- that combination fails at the unpacking."""
    device = next(encoder.parameters()).device
    if isinstance(STFT_batch, np.ndarray):
        STFT_batch = torch.from_numpy(STFT_batch)
    spec = STFT_batch.to(device, torch.float32)
    if quantizer is not None:
        latents = encoder(spec)
        quantized, _loss, (_, _, _) = quantizer(latents)
    else:
        _mu, _logvar, latents = encoder(spec)          # (a tensor [3, ...] would be split along dim 0, as in the reference)
        quantized = None
    imgs, phases, signals = _decode_lists(spec)
    return imgs, phases, signals, latents, quantized
