"""Text-to-timbre pipeline (the hot path of webUI/natural_language_guided_4/text2sound.py:45-179, without the UI):
condition vectors -> DiffSynthSampler.sample (CFG, respaced DDIM/DDPM) -> VectorQuantizerEMA -> Decoder ->
STFT+ decode + iSTFT -> waveforms, and its data-parallel form (prompts sharded over ranks, one all-gather)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, ops, weights as W
from ._lib import check
from .codec import adjust_audio_length, spectrogram_to_waveform, waveform_to_spectrogram
from .sampler import DiffSynthSampler
from .unet import ConditionedUnet
from .vqgan import VQGAN


@dataclass
class Timbres:
    latents: torch.Tensor        # [B,4,128,W]     final continuous latent (imgs[-1])
    quantized: torch.Tensor      # [B,4,128,W]
    spectrograms: torch.Tensor   # [B,3,512,4W]
    waveforms: torch.Tensor      # [B, 256*(4W-1)]


class _Tail:
    """quantiser -> decoder -> STFT+ decode + iSTFT (text2sound.py:128-134, utils.py:224-241) as a fixed launch sequence on
    static buffers, so that it is captured in the same CUDA graph as the sampling loop: ``src`` is the loop's final-latent
    buffer; the quantiser writes straight into the decoder plan's input, the iSTFT reads the decoder plan's output."""

    def __init__(self, vqgan: VQGAN, src: torch.Tensor):
        B, Cc, H, Wd = src.shape
        dev = src.device
        lib = _lib.load()
        self.src = src
        self.dec = vqgan._decoder._stack.plan(B, H, Wd)
        self.q = self.dec.inp                                   # [B,4,H,W] fp32: decoder input = quantiser output
        self.idx = torch.empty((B * H * Wd,), dtype=torch.long, device=dev)
        self.spec = self.dec.out_f32                            # [B,3,4H,4W]
        T = self.spec.shape[3]
        self.frames = torch.empty((B, T, 1024), dtype=torch.float32, device=dev)
        self.wave = torch.empty((B, lib.ds_istft_length(T)), dtype=torch.float32, device=dev)
        vq = vqgan._vq_vae
        K = vq._num_embeddings
        self.ops = [("vq", lambda: vq.quantize_into(self.src, self.q, self.idx))] + list(self.dec.ops) + [
            ("istft", lambda: check(lib.ds_stft_decode_istft(self.spec.data_ptr(), self.frames.data_ptr(), self.wave.data_ptr(), B, T,
                                                             ops._stream()), "ds_stft_decode_istft"))]
        self.meta = dict(self.dec.meta)
        # the quantiser is fp32-ALU work: 4 FMAs per (position, code) for the dot product + the compare (VQGAN.py:107-112)
        self.meta["vq"] = dict(family="vq_quantize", flops=0.0, bytes=B * H * Wd * (2 * Cc * 4.0 + 8.0), fp32_flops=B * H * Wd * K * 2.0 * Cc)
        self.meta["istft"] = dict(family="istft(decode+frames+ola)", flops=0.0, bytes=B * (3 * 512 * T * 4.0 + self.wave.shape[1] * 4.0))
        self.num_launches = len(self.ops) + 1 + sum(1 for n, _ in self.ops if n.endswith(".fin"))      # iSTFT = 2 kernels, finalize = 2

    def run(self):
        for _, fn in self.ops:
            fn()


class _TailBuffers:
    """Result buffers of the tail when it runs inside the library's sampling graph (ds_sample_graph_build): the same attributes
    _Tail exposes (q, idx, spec, wave), no launch list."""

    def __init__(self, src: torch.Tensor):
        B, Cc, H, Wd = src.shape
        dev = src.device
        self.src = src
        self.q = torch.empty((B, Cc, H, Wd), dtype=torch.float32, device=dev)
        self.idx = torch.empty((B * H * Wd,), dtype=torch.long, device=dev)
        self.spec = torch.empty((B, 3, 4 * H, 4 * Wd), dtype=torch.float32, device=dev)
        self.wave = torch.empty((B, _lib.load().ds_istft_length(4 * Wd)), dtype=torch.float32, device=dev)


class _TailFactory:
    """What the sampler needs to put the tail into its graph: the VQGAN (its module-level handle for the library's graph), the
    result buffers, or -- for the operator-level path and bench.py's per-kernel timings -- the Python launch list."""

    def __init__(self, vqgan: VQGAN):
        self.vqgan = vqgan

    def __call__(self, src: torch.Tensor) -> _Tail:
        return _Tail(self.vqgan, src)

    def buffers(self, src: torch.Tensor) -> _TailBuffers:
        return _TailBuffers(src)


class TextToTimbre:
    def __init__(self, unet: ConditionedUnet, vqgan: VQGAN, timesteps: int = 1000, height: int = 128, channels: int = 4,
                 noise_strategy: str = "repeat", device=None):
        self.unet, self.vqgan = unet, vqgan
        self.timesteps, self.height, self.channels, self.noise_strategy = timesteps, height, channels, noise_strategy
        self.device = torch.device(device if device is not None else unet.device)
        self._samplers = {}
        self._tail_factory = _TailFactory(self.vqgan)                # one object: its id() is part of the samplers' graph keys
        self.last_launches = 0
        self._last_tail: Optional[_Tail] = None

    def tail_for(self, batch: int, width: int) -> _Tail:
        """The tail of the most recent generate() of this batch / width as an operator-level launch list on the same final-latent
        buffer (bench.py times its launches one by one)."""
        t = self._last_tail
        assert t is not None and t.src.shape[0] == batch and t.src.shape[3] == width, "no generate() call with this shape yet"
        return t if isinstance(t, _Tail) else _Tail(self.vqgan, t.src)

    def _decode(self, s: DiffSynthSampler, latents: torch.Tensor) -> Timbres:
        """Results of the tail that ran inside the sampler's graph (copied out of its static buffers)."""
        t = s.last_tail
        self._last_tail = t
        self.last_launches = s.last_graph_launches
        self.vqgan._vq_vae.last_indices = t.idx.clone()
        return Timbres(latents, t.q.clone(), t.spec.clone(), t.wave.clone())

    @classmethod
    def random_init(cls, device="cuda", seed: int = 0, perturb_norm: bool = False, batch_invariant: bool = False) -> "TextToTimbre":
        """Deployed architecture (app.py:32-40) with deterministic synthetic weights (no checkpoints offline)."""
        unet = ConditionedUnet(**{k: v for k, v in W.UNET_DEPLOYED.items() if k not in ("out_dim", "time_dim")}, device=device,
                               batch_invariant=batch_invariant)
        unet.load_state_dict(W.unet_random_state_dict(seed=seed, perturb_norm=perturb_norm))
        vq = VQGAN(**W.VQGAN_DEPLOYED, device=device, batch_invariant=batch_invariant)
        vq.load_state_dict(W.vqgan_random_state_dict(seed=seed + 1, perturb_norm=perturb_norm))
        return cls(unet, vq, device=device)

    def sampler_for(self, batch: int, steps: int, cfg_scale: float, uncond: Optional[torch.Tensor]) -> DiffSynthSampler:
        """text2sound.py:96-106: fresh sampler, CFG activated, respaced to ``steps`` (cached per (batch, steps))."""
        key = (batch, steps)
        s = self._samplers.get(key)
        if s is None:
            s = DiffSynthSampler(self.timesteps, height=self.height, channels=self.channels, noise_strategy=self.noise_strategy,
                                 mute=True, device=str(self.device), max_batchsize=batch)
            s.respace(list(np.linspace(0, self.timesteps - 1, steps, dtype=np.int32)))
            self._samplers[key] = s
        s.activate_classifier_free_guidance(cfg_scale, uncond)
        return s

    @torch.no_grad()
    def generate_from_tokens(self, text_encoder, input_ids: torch.Tensor, attention_mask: torch.Tensor, negative_ids: torch.Tensor,
                             negative_mask: torch.Tensor, **kw) -> Timbres:
        """text2sound.py:89-134 with the text front end on the device: ``text_encoder`` (diffusynth_b200.TextEncoder) turns the
        tokenizer's output for B DISTINCT prompts and for the one negative prompt into the condition / unconditional vectors
        (the reference encodes a single prompt on the CPU and repeats it, :89-91,109)."""
        cond = text_encoder.get_text_features(input_ids, attention_mask)
        uncond = text_encoder.get_text_features(negative_ids[:1], negative_mask[:1])[0]
        return self.generate(cond, uncond, **kw)

    @torch.no_grad()
    def generate(self, cond: torch.Tensor, uncond: Optional[torch.Tensor], steps: int = 20, cfg_scale: float = 6, width: int = 64,
                 sampler: str = "ddim", seed: Optional[int] = None, noise_feed: Optional[torch.Tensor] = None,
                 decode: bool = True) -> Timbres:
        B = cond.shape[0]
        s = self.sampler_for(B, steps, cfg_scale, uncond)
        s.noise_feed = noise_feed
        init = None
        if noise_feed is not None:
            init = noise_feed[0][:B].to(self.device, torch.float32)
            s.noise_feed = noise_feed[1:]
        s.graph_tail, s.final_only = (self._tail_factory if decode else None), True
        imgs, _ = s.sample(self.unet, (B, self.channels, self.height, width), return_tensor=True, condition=cond.to(self.device),
                           sampler=sampler, initial_noise=init, seed=seed)
        latents = imgs[-1]
        if not decode:
            self.last_launches = s.last_graph_launches
            return Timbres(latents, None, None, None)
        return self._decode(s, latents)       # quantiser (text2sound.py:128), decoder (utils.py:224), iSTFT (utils.py:229-241): same graph


    # ---- timbre modification (sound2sound_with_text.py:47-269) -------------------------------------------------------
    @torch.no_grad()
    def encode_audio(self, wave: torch.Tensor, width: int = 64) -> torch.Tensor:
        """receive_upload_origin_audio's tensor path (:75-107): peak-normalise, crop/zero-pad to 256*(4w-1) samples, STFT
        (1024/256, hann) -> pad_STFT -> encode_stft -> VQGAN encoder.  Returns the UN-quantised latent [B,4,128,w] (:107)."""
        wave = wave.to(self.device, torch.float32)
        wave = wave / wave.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12)              # :75
        wave = adjust_audio_length(wave, 256 * (4 * width - 1))                            # :80-82
        # pad_STFT(D) is called with its default time_resolution = 256 frames and never crops (tools.py:170-182): clips of
        # <= 3 s (width <= 64) always encode to a 64-column latent, longer ones to ``width`` columns
        spec = waveform_to_spectrogram(wave, time_resolution=max(256, 4 * width))          # :85-94
        return self.vqgan._encoder(spec)

    @torch.no_grad()
    def modify(self, guide_latent: torch.Tensor, cond: torch.Tensor, uncond: Optional[torch.Tensor], steps: int = 20,
               strength: float = 0.7, cfg_scale: float = 6, sampler: str = "ddim", seed: Optional[int] = None,
               noise_feed: Optional[torch.Tensor] = None, decode: bool = True) -> Timbres:
        """sound2sound_sample (:126-269): respace to int(steps/strength) (:185), img_guided_sample(noising_strength=strength,
        guide_img=latent.repeat(B)) (:194-203), then the same VQ -> decoder -> iSTFT tail as text-to-timbre."""
        B = cond.shape[0]
        width = guide_latent.shape[-1]
        n_steps = int(steps / strength)
        key = ("modify", B, n_steps)
        s = self._samplers.get(key)
        if s is None:
            s = DiffSynthSampler(self.timesteps, height=self.height, channels=self.channels, noise_strategy=self.noise_strategy,
                                 mute=True, device=str(self.device), max_batchsize=B)
            s.respace(list(np.linspace(0, self.timesteps - 1, n_steps, dtype=np.int32)))
            self._samplers[key] = s
        s.activate_classifier_free_guidance(cfg_scale, uncond)
        guide = guide_latent.to(self.device, torch.float32)
        if guide.shape[0] == 1 and B > 1:
            guide = guide.repeat(B, 1, 1, 1)
        init = None
        s.noise_feed = None
        if noise_feed is not None:
            init = noise_feed[0][:B].to(self.device, torch.float32)
            s.noise_feed = noise_feed[1:]
        s.graph_tail, s.final_only = (self._tail_factory if decode else None), True
        imgs, _ = s.img_guided_sample(self.unet, (B, self.channels, self.height, width), strength, guide, return_tensor=True,
                                      condition=cond.to(self.device), sampler=sampler, initial_noise=init, seed=seed)
        latents = imgs[-1]
        if not decode:
            self.last_launches = s.last_graph_launches
            return Timbres(latents, None, None, None)
        return self._decode(s, latents)


    # ---- per-note synthesis for arrangements (track_maker.py:228-283) ---------------------------------------------------
    @staticmethod
    def note_width(duration_sec: float, time_resolution: int = 256, vae_scale: int = 4) -> int:
        return int(time_resolution * ((duration_sec + 1) / 4) / vae_scale)                                    # track_maker.py:245

    @torch.no_grad()
    def _synthesize_group(self, instrument_latent, condition, width: int, B: int, sample_steps, noising_strength, attack, before_release,
                          sampler, time_resolution, vae_scale, noise_feed) -> Timbres:
        """B notes of one width as ONE batched inpainting run (one graph launch): every note is an independent sample of the batch
        (own noise rows), exactly what B separate calls of the reference's per-note closure compute."""
        # the reference builds a fresh sampler per note (:248-249); here it is cached per (step count, batch) so that the captured
        # graph of a (width, steps, batch) triple is reused by every later group of that shape instead of being re-captured
        key = ("note", sample_steps, B)
        s = self._samplers.get(key)
        if s is None:
            s = DiffSynthSampler(self.timesteps, height=self.height, channels=self.channels, noise_strategy="repeat", mute=True,
                                 device=str(self.device), max_batchsize=B)                                     # :248
            s.respace(list(np.linspace(0, self.timesteps - 1, sample_steps, dtype=np.int32)))                  # :249
            self._samplers[key] = s
        s.activate_classifier_free_guidance(1.0, None)
        s.noise_feed = None
        s.graph_tail, s.final_only = self._tail_factory, True
        mask = torch.zeros((B, 1, self.height, width), dtype=torch.float32, device=self.device)                # :252-254
        mask[:, :, :, :int(time_resolution * (attack / 4) / vae_scale)] = 1.0
        mask[:, :, :, -int(time_resolution * ((before_release + 1) / 4) / vae_scale):] = 1.0
        init = None
        if noise_feed is not None:
            init = noise_feed[0][:B].to(self.device, torch.float32)
            s.noise_feed = noise_feed[1:]
        guide = instrument_latent.to(self.device, torch.float32)
        if guide.shape[0] == 1 and B > 1:
            guide = guide.repeat(B, 1, 1, 1)
        cond = condition.to(self.device)
        if cond.shape[0] == 1 and B > 1:
            cond = cond.repeat(B, 1)
        imgs, _ = s.inpaint_sample(self.unet, (B, self.channels, self.height, width), noising_strength, guide, mask, return_tensor=True,
                                   condition=cond, sampler=sampler, initial_noise=init,
                                   use_dynamic_mask=True, end_noise_level_ratio=0.0, mask_flexivity=1.0)       # :257-269
        return self._decode(s, imgs[-1])       # quantiser :275, decoder + iSTFT :277-283, inside the group's graph

    @torch.no_grad()
    def synthesize_note(self, instrument_latent: torch.Tensor, condition: torch.Tensor, duration_sec: float, sample_steps: int = 20,
                        noising_strength: float = 1.0, attack: float = 0.5, before_release: float = 0.5, sampler: str = "ddim",
                        time_resolution: int = 256, vae_scale: int = 4, noise_feed: Optional[torch.Tensor] = None) -> Timbres:
        """One note of an instrument at an arbitrary duration: the ``diffSynthSampler`` closure of track_maker.DiffSynth.
        ``instrument_latent`` [1,4,128,64] is the instrument's latent (re-laid to the note's width by the repeat strategy),
        ``condition`` [1,512] the empty-prompt embedding the reference passes (:231-233; no classifier-free guidance here);
        attack and tail are frozen by a mask that shrinks over the steps (``use_dynamic_mask``, mask_flexivity 1.0).
        The width ``int(256 * ((duration + 1) / 4) / 4)`` is arbitrary, odd levels included (pad_to_match)."""
        width = self.note_width(duration_sec, time_resolution, vae_scale)
        return self._synthesize_group(instrument_latent, condition, width, 1, sample_steps, noising_strength, attack, before_release,
                                      sampler, time_resolution, vae_scale, noise_feed)

    @torch.no_grad()
    def synthesize_notes(self, instrument_latent: torch.Tensor, condition: torch.Tensor, durations_sec, sample_steps: int = 20,
                         noising_strength: float = 1.0, attack: float = 0.5, before_release: float = 0.5, sampler: str = "ddim",
                         time_resolution: int = 256, vae_scale: int = 4, noise_feeds=None) -> list:
        """The notes of a track (track_maker.py:285-330 calls the per-note closure once per distinct duration): notes of equal
        duration -- hence equal latent width -- are batched into ONE graph launch.  Returns one ``Timbres`` per
        note, in input order.  ``noise_feeds`` (tests): per note a host noise tensor [count, 1, C, H, 64]."""
        widths = [self.note_width(d, time_resolution, vae_scale) for d in durations_sec]
        out = [None] * len(widths)
        for w in sorted(set(widths)):
            idx = [i for i, ww in enumerate(widths) if ww == w]
            feed = torch.cat([noise_feeds[i] for i in idx], dim=1) if noise_feeds is not None else None
            t = self._synthesize_group(instrument_latent, condition, w, len(idx), sample_steps, noising_strength, attack, before_release,
                                       sampler, time_resolution, vae_scale, feed)
            for b, i in enumerate(idx):
                out[i] = Timbres(t.latents[b:b + 1], t.quantized[b:b + 1], t.spectrograms[b:b + 1], t.waveforms[b:b + 1])
        return out


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of ``total`` prompts: rank r owns [lo, hi); per = ceil(total/world)."""
    per = -(-total // world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def all_gather_waveforms(local: torch.Tensor, total: int, world: int) -> torch.Tensor:
    """One equal-count all-gather of [per, L] waveforms (NCCL on GPUs, gloo in CPU tests) -> [total, L] on every rank."""
    import torch.distributed as dist
    per = -(-total // world)
    if local.shape[0] < per:
        local = torch.cat([local, local.new_zeros((per - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:total]


class ChunkGatherer:
    """The collective half of the sharded job: rank-local result chunks [chunk, L] are all-gathered one by one (NCCL on a side
    stream on GPUs, so that chunk k travels while chunk k+1 is computed; gloo / plain copies on CPU tensors in the host tests) into
    the rank-major result [world, per_pad, L]; ``finish()`` returns the first ``total`` rows in prompt order."""

    def __init__(self, total: int, rank: int, world: int, chunk: int, L: int, device, dtype=torch.float32, comm=None):
        self.total, self.rank, self.world, self.chunk, self.L = total, rank, world, chunk, L
        self.comm_handle = comm      # engine.Comm (the library's own NCCL communicator, ds_allgather) or None = torch.distributed
        self.per = -(-total // world)
        self.n_chunks = -(-self.per // chunk)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.out = torch.zeros((world, self.n_chunks * chunk, L), dtype=dtype, device=self.device)
        self.stage = [torch.empty((world, chunk, L), dtype=dtype, device=self.device) for _ in range(2)]
        self._exposed = None

    def submit(self, k: int, w: torch.Tensor) -> None:
        """Chunk k of this rank (rows past the rank's share zero-padded by the caller to [chunk, L])."""
        import torch.distributed as dist
        assert tuple(w.shape) == (self.chunk, self.L), tuple(w.shape)

        def gather():
            if self.world > 1:
                st = self.stage[k % 2]
                if self.comm_handle is not None:
                    self.comm_handle.all_gather(w, recv=st.view(self.world * self.chunk, self.L))
                else:
                    dist.all_gather_into_tensor(st.view(self.world * self.chunk, self.L), w.contiguous())
                self.out[:, k * self.chunk:(k + 1) * self.chunk].copy_(st)
            else:
                self.out[0, k * self.chunk:(k + 1) * self.chunk].copy_(w)

        if not self.cuda:
            gather()
            return
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            gather()
            w.record_stream(self.comm)

    def finish(self) -> torch.Tensor:
        if self.cuda:
            main = torch.cuda.current_stream(self.device)
            t_main, t_comm = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_main.record(main)
            t_comm.record(self.comm)
            main.wait_stream(self.comm)
            self._exposed = (t_main, t_comm)
        return self.out[:, :self.per].reshape(self.world * self.per, self.L)[:self.total]

    def exposed_ms(self) -> float:
        """How long the side stream ran past the end of the last chunk's compute in the most recent job (synchronises)."""
        if self._exposed is None:
            return 0.0
        t_main, t_comm = self._exposed
        t_comm.synchronize()
        return max(0.0, t_main.elapsed_time(t_comm))


class ShardedGenerator:
    """BASELINE configs[3] / SURVEY 8(e): ``total`` prompts split contiguously over ``world`` ranks; each rank samples its share in
    chunks of ``chunk`` prompts through the text-to-timbre graph and the waveforms of chunk k are all-gathered on a side stream
    while chunk k+1 is being sampled, so that only the last chunk's gather is exposed.  Every rank ends with all waveforms
    [total, L] in prompt order."""

    def __init__(self, pipe: TextToTimbre, total: int, rank: int, world: int, chunk: int = 64, comm=None):
        self.pipe, self.total, self.rank, self.world, self.chunk = pipe, total, rank, world, chunk
        self.comm = comm
        self.lo, self.hi = shard_range(total, rank, world)
        self.gatherer: Optional[ChunkGatherer] = None
        self.n_chunks = -(-(-(-total // world)) // chunk)
        self.launches_per_job = 0

    @torch.no_grad()
    def run(self, cond: torch.Tensor, uncond: Optional[torch.Tensor], steps: int = 20, cfg_scale: float = 6, width: int = 64,
            sampler: str = "ddim", seed: Optional[int] = None) -> torch.Tensor:
        pipe, chunk = self.pipe, self.chunk
        L = 256 * (4 * width - 1)
        if self.gatherer is None or self.gatherer.L != L:
            self.gatherer = ChunkGatherer(self.total, self.rank, self.world, chunk, L, pipe.device, comm=self.comm)
        n_local = self.hi - self.lo
        launches = 0
        for k in range(self.n_chunks):
            c = cond[k * chunk:min((k + 1) * chunk, n_local)]
            if c.shape[0] > 0:
                w = pipe.generate(c, uncond, steps=steps, cfg_scale=cfg_scale, width=width, sampler=sampler, seed=seed).waveforms
                launches += pipe.last_launches
                if w.shape[0] < chunk:
                    w = torch.cat([w, w.new_zeros((chunk - w.shape[0], L))])
            else:
                w = torch.zeros((chunk, L), dtype=torch.float32, device=pipe.device)       # (a rank past the end of the prompt list)
            self.gatherer.submit(k, w)
        self.launches_per_job = launches
        return self.gatherer.finish()

    def exposed_gather_ms(self) -> float:
        return self.gatherer.exposed_ms() if self.gatherer is not None else 0.0
