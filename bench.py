#!/usr/bin/env python
"""bench.py -- timbres/sec of the text-to-timbre path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch of synthetic prompts: CFG-doubled U-Net x `sample_steps`
DDIM steps (one CUDA graph) -> VectorQuantizerEMA -> VQGAN decoder -> STFT+ decode + iSTFT -> waveforms
(BASELINE.json configs[2] = configs[1] + decode), batch 64 per GPU, deployed architecture, random-init weights.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
N>1: launched by torchrun, one rank per GPU; prompts are sharded (weak scaling: 64 per GPU), each rank samples
independently and the waveforms are all-gathered over NCCL inside the timed region.

--impl reference times the reference's CPU implementation of the same path (the oracle port of it: the reference
tree does not exist on the GPU box) on the host cores, each step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
WIDTH = 64
CFG_SCALE = 6
METRIC = "timbres_per_sec"
SHARDED_TOTAL = 1024
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal CUDA-core fp32 FMA peak of a B200 at its 1965 MHz maximum clock (74.4)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_models():
    from diffusynth_b200 import weights as W
    usd, vsd = W.unet_random_state_dict(seed=0, perturb_norm=False), W.vqgan_random_state_dict(seed=1, perturb_norm=False)
    _, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    return usd, vsd, dec_plan


def cpu_timbre(models, sample_steps: int, noise_seed: int = 0):
    """ONE full batch-1 timbre on the host: `sample_steps` CFG-doubled U-Net steps -> quantiser -> decoder -> iSTFT (BASELINE.json
    configs[0] / BASELINE.md section 3), fp32 torch oracle port of the reference.  Returns (timings, latent, waveform)."""
    from diffusynth_b200 import weights as W
    from oracle import ds_oracle as O
    usd, vsd, dec_plan = models
    cond, uncond = W.synthetic_conditions(1, 512)
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, sample_steps, dtype=np.int32)))
    draws = W.host_noise(noise_seed, 1 + sample_steps, 1)
    with torch.no_grad():
        t0 = time.perf_counter()
        lat = O.sample_loop(lambda x, t, c: O.unet_forward(usd, x, t, c), sch, (1, 4, 128, WIDTH), cond, uncond, CFG_SCALE, draws)[-1]
        t1 = time.perf_counter()
        # quantiser timed the way the reference runs it: torch fp32 distance matrix + argmin (VQGAN.py:107-117)
        flat = lat.permute(0, 2, 3, 1).reshape(-1, 4)
        cb = vsd["_vq_vae._embedding.weight"]
        d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(cb ** 2, dim=1) - 2 * torch.matmul(flat, cb.t())
        idx = torch.argmin(d, dim=1)
        q = cb[idx].view(1, 128, WIDTH, 4).permute(0, 3, 1, 2).contiguous()
        t2 = time.perf_counter()
        spec = O.vqgan_decode(vsd, dec_plan, q)
        t3 = time.perf_counter()
        wave = O.spectrogram_to_waveform(spec[0].numpy().astype(np.float64))
        t4 = time.perf_counter()
    return dict(total=t4 - t0, sampler=t1 - t0, vq=t2 - t1, decoder=t3 - t2, istft=t4 - t3), lat, wave


def cpu_arm(sample_steps: int, threads: int | None = None, repeats: int = 3, also_steps=(10,), keep_outputs: bool = False):
    """batch 1, fp32, all host cores, 1 warm-up, median of `repeats` FULL runs (no extrapolation) at the
    benchmark's step count, plus the same at the reference's CPU default of 10 steps."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    models = _cpu_models()
    cpu_timbre(models, 2)                                   # warm-up (thread pool, allocator, oneDNN primitive caches)
    res = {}
    outs = None
    for st in tuple(also_steps) + (sample_steps,):
        runs = []
        for _ in range(repeats):
            tm, lat, wave = cpu_timbre(models, st)
            runs.append(tm)
            if st == sample_steps:
                outs = (lat, wave)
        runs.sort(key=lambda r: r["total"])
        res[st] = runs[len(runs) // 2]
    m = res[sample_steps]
    sample = (f"batch 1, median of {repeats} full runs after 1 warm-up: {sample_steps} CFG-doubled U-Net steps {m['sampler']:.2f} s "
              f"({m['sampler'] / sample_steps:.3f} s/step) + VQ {m['vq']:.3f} s + decoder {m['decoder']:.3f} s + iSTFT {m['istft']:.3f} s = {m['total']:.2f} s/timbre; "
              + "; ".join(f"{st} steps: {res[st]['total']:.2f} s/timbre" for st in also_steps)
              + "; fp32 torch oracle port of the reference (the reference tree does not exist on the GPU box)")
    out = dict(value=1.0 / m["total"], unit="timbres/s", cores=threads, kind="port", sample=sample,
               s_per_timbre={str(st): res[st]["total"] for st in res}, s_per_unet_step=m["sampler"] / sample_steps,
               split_s={k: m[k] for k in ("sampler", "vq", "decoder", "istft")})
    if keep_outputs:
        out["_outputs"] = outs
    return out


def run_reference(args):
    """Reference arm: every step is ONE full batch-1 timbre of the workload (all `sample_steps` U-Net steps + tail) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    models = _cpu_models()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_timbre(models, 2)
    t0 = time.perf_counter()
    runs = [cpu_timbre(models, args.sample_steps)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = args.steps / wall
    m = sorted(runs, key=lambda r: r["total"])[len(runs) // 2]
    sample = (f"each step = one full batch-1 timbre ({args.sample_steps} CFG-doubled U-Net steps + VQ + decoder + iSTFT), no extrapolation; median split: "
              f"sampler {m['sampler']:.2f} s, VQ {m['vq']:.3f} s, decoder {m['decoder']:.3f} s, iSTFT {m['istft']:.3f} s; fp32 torch oracle port of the reference")
    line = dict(metric=METRIC, value=v, unit="timbres/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * wall / max(1, args.steps), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(args, args.gpus),
                cpu_baseline=dict(value=v, unit="timbres/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=v, unit="timbres/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(json.dumps(line))


def workload_config(args, world):
    base = dict(batch_per_gpu=BATCH, global_batch=BATCH * world, sample_steps=args.sample_steps, cfg_scale=CFG_SCALE, sampler="ddim",
                operand_dtype="fp16 operands, fp32 accumulate (north_star says bf16; the same tcgen05 kind::f16 instruction and rate: bf16 "
                              "activations fail the 1e-2 gate under CFG 6, DESIGN.md section 3; range evidence: tests/test_gpu_headline.py)",
                l2_policy="per-step working set (multi-GB of activations) exceeds the 126 MB L2; no explicit flush")
    if args.workload == "sharded1024":
        base.update(workload=f"BASELINE configs[3]: {SHARDED_TOTAL} prompts sharded over {world} GPU(s), {SHARDED_TOTAL // world} per rank in chunks of {BATCH} "
                             f"through the text-to-timbre graph ({args.sample_steps} DDIM steps, CFG {CFG_SCALE}, VQ, decoder, iSTFT), chunk k all-gathered on a side "
                             "stream while chunk k+1 samples; every rank ends with all waveforms [1024, 65280] fp32",
                    global_batch=SHARDED_TOTAL, parallelism=f"dp{world} (prompts sharded, chunked all-gather of waveforms overlapped with sampling)")
    elif args.workload == "modify":
        base.update(workload=f"BASELINE configs[4]: timbre modification -- waveform -> STFT+ -> VQGAN encoder -> q_sample(strength 0.7) -> "
                             f"{int(int(args.sample_steps / 0.7) * 0.7)} CFG-doubled U-Net steps (respaced to {int(args.sample_steps / 0.7)}) -> VQ -> decoder -> iSTFT; batch sweep "
                             f"over {world} GPU(s), per-rank share in chunks of <= {BATCH}",
                    parallelism=f"dp{world} (batch sharded, final all-gather of waveforms)")
    else:
        base.update(workload=f"text-to-timbre full pipeline (BASELINE configs[2]): batch {BATCH}/GPU, {args.sample_steps} DDIM steps, CFG {CFG_SCALE}, "
                             f"latent 4x128x{WIDTH} -> VQ(8192) -> decoder [3,512,{4 * WIDTH}] -> iSTFT {256 * (4 * WIDTH - 1)} samples; deployed U-Net (106.9M) + VQGAN; "
                             "sampling loop + tail in ONE CUDA graph",
                    parallelism=f"dp{world} (prompts sharded, final all-gather of waveforms)" if world > 1 else "single GPU")
    return base


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def timed_ops(op_list, reps: int = 2):
    """Eager (un-graphed) pass over a plan's ops with CUDA events around every launch (on torch's current stream, which is
    the stream the launches go to); best of `reps` per op.  -> {op name: ms}."""
    best = {}
    for _ in range(reps):
        evs = []
        torch.cuda.synchronize()
        for name, fn in op_list:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((name, e0, e1))
        torch.cuda.synchronize()
        for name, e0, e1 in evs:
            ms = e0.elapsed_time(e1)
            best[name] = min(best.get(name, 1e30), ms)
    return best


def family_rooflines(ms_by_op, meta, pk, per=1.0):
    """Group ops into kernel families; per family: launches, ms, algorithmic FLOPs / bytes, the roofline time
    max(flops / tensor peak, bytes / HBM peak), which roof binds, and frac = roofline time / measured time."""
    fams = {}
    for name, ms in ms_by_op.items():
        m = meta.get(name, dict(family="misc", flops=0.0, bytes=0.0))
        f = fams.setdefault(m["family"], dict(launches=0, ms=0.0, gflop=0.0, mbytes=0.0, gflop32=0.0))
        f["launches"] += 1
        f["ms"] += ms
        f["gflop"] += m["flops"] / 1e9
        f["mbytes"] += m["bytes"] / 1e6
        f["gflop32"] += m.get("fp32_flops", 0.0) / 1e9
    out = {}
    for k, f in sorted(fams.items(), key=lambda kv: -kv[1]["ms"]):
        t_tensor = f["gflop"] / pk["tf_sustained"]                 # GFLOP / (TFLOP/s) = ms
        t_hbm = f["mbytes"] / pk["hbm"]                            # MB / (GB/s) = ms
        t_fp32 = f["gflop32"] / FP32_PEAK_TFLOPS                   # CUDA-core fp32 work (the quantiser's distance search)
        t_roof = max(t_tensor, t_hbm, t_fp32)
        out[k] = dict(launches=f["launches"], ms=round(f["ms"] / per, 4), algorithmic_gflop=round((f["gflop"] + f["gflop32"]) / per, 2),
                      algorithmic_mbytes=round(f["mbytes"] / per, 1),
                      bound="fp32-alu" if t_fp32 >= max(t_tensor, t_hbm) else "tensor" if t_tensor >= t_hbm else "hbm",
                      achieved_tflops=round(f["gflop"] / f["ms"], 1) if f["ms"] > 0 else None,
                      achieved_gbs=round(f["mbytes"] / f["ms"], 1) if f["ms"] > 0 else None,
                      frac=round(t_roof / f["ms"], 4) if f["ms"] > 0 and t_roof > 0 else None)
    return out


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from diffusynth_b200 import TextToTimbre, weights as W
    from diffusynth_b200.pipeline import all_gather_waveforms, shard_range

    pipe = TextToTimbre.random_init(device=dev, seed=0)
    total = BATCH * world
    lo, hi = shard_range(total, rank, world)
    cond_all, uncond = W.synthetic_conditions(total, 512)
    cond_host = cond_all[lo:hi].contiguous().pin_memory()
    uncond_dev = uncond.to(dev)
    cond_dev = cond_host.to(dev)
    wave_host = torch.empty((BATCH, 256 * (4 * WIDTH - 1)), dtype=torch.float32).pin_memory()
    finite = []

    def step_device():
        out = pipe.generate(cond_dev, uncond_dev, steps=args.sample_steps, cfg_scale=CFG_SCALE, width=WIDTH, seed=None)
        finite.append(torch.isfinite(out.waveforms).all())            # device-side flag, read after the timed region
        if world > 1:
            return all_gather_waveforms(out.waveforms, total, world)
        return out.waveforms

    def step_e2e():
        c = cond_host.to(dev, non_blocking=True)
        out = pipe.generate(c, uncond_dev, steps=args.sample_steps, cfg_scale=CFG_SCALE, width=WIDTH, seed=None)
        w = out.waveforms
        if world > 1:
            w = all_gather_waveforms(w, total, world)[lo:hi]
        wave_host.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return wave_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    torch.manual_seed(1234 + rank)
    warm = max(3, args.warmup)
    for _ in range(warm):
        step_device()
    finite.clear()
    with ClockSampler(local) as clk:
        ms_dev = timed(step_device, args.steps)
    clocks = clk.summary()
    all_finite = all(bool(f) for f in finite)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    all_finite = all_finite and bool(np.isfinite(wave_host.numpy()).all())

    value = total * args.steps / (ms_dev / 1e3)
    e2e_value = total * args.steps / (ms_e2e / 1e3)
    if rank == 0:
        pk = peaks()
        sampler = pipe._samplers[(BATCH, args.sample_steps)]
        loop = next(iter(sampler._graphs.values()))
        plan = loop.plan
        # ---- per-kernel rooflines of one U-Net evaluation (CFG-doubled batch) and of the tail, measured here with CUDA events ----
        unet_ms = timed_ops(plan.ops)
        if os.environ.get("DS_DUMP_OPS"):
            with open(os.environ["DS_DUMP_OPS"], "w") as f:
                json.dump([dict(op=k, ms=round(v, 4), **plan.meta.get(k, {})) for k, v in unet_ms.items()], f, indent=0)
        unet_fams = family_rooflines(unet_ms, plan.meta, pk)
        conv_keys = [k for k in unet_fams if k.startswith("conv")]
        conv_ms = sum(unet_fams[k]["ms"] for k in conv_keys)
        conv_gflop = sum(unet_fams[k]["algorithmic_gflop"] for k in conv_keys)
        conv_launches = sum(unet_fams[k]["launches"] for k in conv_keys)
        achieved = conv_gflop / conv_ms
        tail = pipe.tail_for(BATCH, WIDTH)
        tail_ms = timed_ops(tail.ops)
        tail_fams = family_rooflines(tail_ms, tail.meta, pk)
        # time of the sampling graph alone (U-Net step ms)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            loop.launch()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / 3
        tail_graph_ms = max(0.0, ms_dev / args.steps - graph_ms) if world == 1 else None
        step_ms = (graph_ms - (sum(tail_ms.values()) if loop.has_tail else 0.0)) / loop.n_iter
        algo_step_tflop = 2 * BATCH * 136.70e9 / 1e12
        launches = pipe.last_launches
        # ---- latency of the batch-1 per-note path (track_maker.py:228-283): one 2-second note (width 48), same step count, cached graph ----
        inst = (torch.randn((1, 4, 128, 64), generator=torch.Generator().manual_seed(22)) * 0.9).to(dev)      # synthetic instrument latent
        note = lambda: pipe.synthesize_note(inst, cond_all[:1].to(dev), 2.0, sample_steps=args.sample_steps)
        note(); note()
        lat = []
        for _ in range(5):
            torch.cuda.synchronize()
            e0.record(); note(); e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        note_latency = dict(ms=float(np.median(lat)), batch=1, width=48, sample_steps=args.sample_steps,
                            what="TextToTimbre.synthesize_note: dynamic-mask inpainting + VQ + decoder + iSTFT, one cached graph launch")
        # ---- CPU baseline doubling as the validation reference: the same prompt + host noise on the GPU ----
        cpu, validation = None, dict(finite=all_finite)
        if world == 1:
            cpu = cpu_arm(args.sample_steps, keep_outputs=True)
            lat_cpu, wave_cpu = cpu.pop("_outputs")
            draws = W.host_noise(0, 1 + args.sample_steps, 1)
            out = pipe.generate(cond_all[:1].to(dev), uncond_dev, steps=args.sample_steps, cfg_scale=CFG_SCALE, width=WIDTH, noise_feed=draws)
            rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
            # tail: the quantiser is an argmin (a 1e-3 latent difference flips a few indices), so -- like the parity tests -- the tail is
            # judged on IDENTICAL quantiser inputs: oracle tail on the GPU's latents; the free-running end-to-end figure is reported too
            from oracle import ds_oracle as O
            usd_, vsd_, dec_plan_ = _cpu_models()
            with torch.no_grad():
                q_same, idx_same = O.vq_quantize(out.latents.cpu(), vsd_["_vq_vae._embedding.weight"])
                wave_same = O.spectrogram_to_waveform(O.vqgan_decode(vsd_, dec_plan_, q_same)[0].numpy().astype(np.float64))
                _, idx_cpu = O.vq_quantize(lat_cpu, vsd_["_vq_vae._embedding.weight"])
            validation.update(latent_rel_l2=rel(out.latents, lat_cpu), vq_bit_exact=bool(torch.equal(q_same, out.quantized.cpu())),
                              waveform_rel_l2=rel(out.waveforms[0], torch.from_numpy(wave_same)),
                              waveform_rel_l2_free_running=rel(out.waveforms[0], torch.from_numpy(wave_cpu)),
                              vq_index_flips_free_running=float((idx_same != idx_cpu).float().mean()),
                              against=f"CPU oracle port, same prompt and host noise, batch 1, {args.sample_steps} steps, CFG {CFG_SCALE}; tail (VQ, decoder, "
                                      "iSTFT) on the GPU's own final latents",
                              tolerance=dict(latent=1e-2, waveform=2e-2))
            validation["passed"] = bool(all_finite and validation["latent_rel_l2"] < 1e-2 and validation["vq_bit_exact"]
                                        and validation["waveform_rel_l2"] < 2e-2)
        else:
            validation["passed"] = bool(all_finite)
        line = dict(metric=METRIC, value=value, unit="timbres/s", n_gpus=world, steps=args.steps, warmup=warm,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp16", data="synthetic",
                    config=workload_config(args, world), clocks=clocks,
                    e2e=dict(value=e2e_value, unit="timbres/s", h2d_bytes_per_step=int(cond_host.numel() * 4),
                             d2h_bytes_per_step=int(wave_host.numel() * 4), ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(launches * args.steps),
                    validated=validation["passed"], validation=validation,
                    unet_step_ms=step_ms,
                    unet_step_roofline=dict(algorithmic_tflop=algo_step_tflop, achieved_tflops=algo_step_tflop / (step_ms / 1e3),
                                            peak_tflops=pk["tf_sustained"], frac=algo_step_tflop / (step_ms / 1e3) / pk["tf_sustained"]),
                    roofline=dict(kernel="conv_gemm_kernel (tcgen05 implicit-GEMM conv; all dense convs of the U-Net)", bound="tensor",
                                  achieved=achieved, peak=pk["tf_sustained"], unit="TFLOP/s", frac=achieved / pk["tf_sustained"], traffic=conv_traffic(),
                                  traffic_note="ncu dram__bytes_read+write per conv_gemm launch, averaged over the conv launches of one U-Net evaluation "
                                               "(committed capture under profiles/; not re-measured inside this run)",
                                  peak_source=pk["source"] + " bf16_tflops_sustained", launches_per_unet_eval=conv_launches,
                                  algorithmic_gflop_per_unet_eval=conv_gflop, ms_per_unet_eval=conv_ms),
                    roofline_per_kernel=dict(
                        note="CUDA-event time of every launch of one eager U-Net evaluation (128 guidance-doubled samples) and one eager tail pass "
                             "(64 timbres); frac = max(algorithmic FLOPs / sustained bf16 peak, algorithmic bytes / HBM copy peak) / measured time",
                        peaks=dict(tensor_tflops=pk["tf_sustained"], hbm_gbs=pk["hbm"], source=pk["source"]),
                        unet_eval=unet_fams, tail=tail_fams),
                    unet_eval_ms_eager=sum(unet_ms.values()), tail_ms_eager=sum(tail_ms.values()), sampling_graph_ms=graph_ms,
                    note_latency=note_latency)
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "s_per_timbre", "s_per_unet_step", "split_s")}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local, dev


def _timed_job(dist, world, dev, fn, K):
    """K passes of fn between barriers, CUDA events on the current stream, max over ranks -> ms per pass."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / K


def run_sharded1024(args):
    """BASELINE configs[3] (strong scaling): 1024 prompts, contiguous shards, chunks of 64 per rank; the all-gather of chunk k runs on
    a side stream while chunk k+1 is sampled."""
    dist, world, rank, local, dev = _dist_setup()
    from diffusynth_b200 import TextToTimbre, weights as W
    from diffusynth_b200.pipeline import ShardedGenerator
    pipe = TextToTimbre.random_init(device=dev, seed=0)
    cond_all, uncond = W.synthetic_conditions(SHARDED_TOTAL, 512)
    comm = None
    if world > 1:          # the job's one collective runs on the library's own communicator (ds_comm_init / ds_allgather)
        from diffusynth_b200.engine import Comm

        def bootstrap(ident):
            buf = ident.to(dev)
            dist.broadcast(buf, src=0)
            return buf.cpu()
        comm = Comm(rank, world, bootstrap)
    gen = ShardedGenerator(pipe, SHARDED_TOTAL, rank, world, chunk=BATCH, comm=comm)
    cond_host = cond_all[gen.lo:gen.hi].contiguous().pin_memory()
    uncond_dev = uncond.to(dev)
    result_host = torch.empty((SHARDED_TOTAL, 256 * (4 * WIDTH - 1)), dtype=torch.float32).pin_memory() if rank == 0 else None

    def job():
        c = cond_host.to(dev, non_blocking=True)                      # host -> device inside the timed region
        out = gen.run(c, uncond_dev, steps=args.sample_steps, cfg_scale=CFG_SCALE, width=WIDTH)
        if rank == 0:
            result_host.copy_(out, non_blocking=True)                 # device -> host of the gathered result
        torch.cuda.current_stream().synchronize()
        return out

    torch.manual_seed(1234 + rank)
    job()
    job()                                                            # W: graph capture + NCCL channel setup happen here
    with ClockSampler(local) as clk:
        ms = _timed_job(dist, world, dev, job, args.steps)
    out = job()
    ok = bool(torch.isfinite(out).all()) and tuple(out.shape) == (SHARDED_TOTAL, 256 * (4 * WIDTH - 1))
    if rank == 0:
        v = SHARDED_TOTAL / (ms / 1e3)
        line = dict(metric=METRIC, value=v, unit="timbres/s", n_gpus=world, steps=args.steps, warmup=2, ms_per_step=ms, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="fp16", data="synthetic", config=workload_config(args, world), clocks=clk.summary(),
                    e2e=dict(value=v, unit="timbres/s", h2d_bytes_per_step=int(cond_host.numel() * 4), d2h_bytes_per_step=int(result_host.numel() * 4), ms_per_step=ms),
                    gpu_launches=int(gen.launches_per_job * args.steps), validated=ok,
                    validation=dict(finite=ok, note="output checked finite and complete; parity of this path: tests/test_gpu_headline.py, tests/test_gpu_fullsize.py"),
                    chunks_per_rank=gen.n_chunks, gather=dict(bytes_total=int(SHARDED_TOTAL * 256 * (4 * WIDTH - 1) * 4), overlapped=world > 1, collective="ds_allgather (library NCCL communicator)" if comm is not None else "none (1 rank)",
                                                             ms_exposed_after_last_chunk=gen.exposed_gather_ms()))
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_modify(args):
    """BASELINE configs[4]: timbre modification, batch sweep 1 .. max over the available GPUs."""
    dist, world, rank, local, dev = _dist_setup()
    from diffusynth_b200 import TextToTimbre, weights as W
    from diffusynth_b200.pipeline import all_gather_waveforms, shard_range
    from oracle import cases                      # (synthetic input clip only)
    pipe = TextToTimbre.random_init(device=dev, seed=0)
    wave_host = torch.from_numpy(cases.synthetic_wave(seed=33)).float()[None].pin_memory()
    max_b = args.max_batch or (4096 if world >= 8 else 1024)
    sweep, clocks = [], None
    b = 1
    while b <= max_b:
        lo, hi = shard_range(b, rank, world)
        cond_all, uncond = W.synthetic_conditions(b, 512)
        cond_host = cond_all[lo:hi].contiguous().pin_memory() if hi > lo else None
        uncond_dev = uncond.to(dev)

        def job():
            waves = []
            if hi > lo:
                guide = pipe.encode_audio(wave_host.to(dev, non_blocking=True))             # STFT+ -> encoder, once per job
                c = cond_host.to(dev, non_blocking=True)
                for c0 in range(0, hi - lo, BATCH):
                    cc = c[c0:c0 + BATCH]
                    waves.append(pipe.modify(guide, cc, uncond_dev, steps=args.sample_steps, strength=0.7, cfg_scale=CFG_SCALE).waveforms)
            w = torch.cat(waves) if waves else torch.zeros((0, 256 * (4 * WIDTH - 1)), device=dev)
            if world > 1:
                w = all_gather_waveforms(w, b, world)
            torch.cuda.current_stream().synchronize()
            return w

        torch.manual_seed(99 + rank)
        job()
        job()
        if b == max_b:
            with ClockSampler(local) as clk:
                ms = _timed_job(dist, world, dev, job, args.steps)
            clocks = clk.summary()
        else:
            ms = _timed_job(dist, world, dev, job, args.steps)
        w = job()
        ok = bool(torch.isfinite(w).all()) and w.shape[0] == b
        sweep.append(dict(batch=b, ms=round(ms, 3), timbres_per_sec=round(b / (ms / 1e3), 2), finite=ok))
        b *= 2
    if rank == 0:
        top = sweep[-1]
        line = dict(metric=METRIC, value=top["timbres_per_sec"], unit="timbres/s", n_gpus=world, steps=args.steps, warmup=2, ms_per_step=top["ms"],
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="fp16", data="synthetic", config=workload_config(args, world),
                    clocks=clocks, e2e=dict(value=top["timbres_per_sec"], unit="timbres/s", h2d_bytes_per_step=int(wave_host.numel() * 4 + top["batch"] * 512 * 4 // world),
                                            d2h_bytes_per_step=0, ms_per_step=top["ms"]),
                    gpu_launches=int(pipe.last_launches * args.steps), validated=all(r["finite"] for r in sweep),
                    validation=dict(finite=all(r["finite"] for r in sweep), note="parity of this path: tests/test_gpu_headline.py::test_headline_config_parity[modify20]"),
                    sweep=sweep, note="value = throughput at the largest batch of the sweep; latency at batch 1 = sweep[0].ms")
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Native libraries (NCCL's version banner) write to fd 1; stdout must carry exactly one JSON line, so fd 1 is pointed at
    stderr for the whole run and the line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def conv_traffic():
    """DRAM bytes per conv_gemm launch from the committed ncu capture (profiles/), or None when the capture is absent."""
    for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", name)
        try:
            with open(path) as f:
                return float(json.load(f)["avg_dram_bytes_per_conv_launch"])
        except (OSError, KeyError, ValueError):
            continue
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sample-steps", type=int, default=20, dest="sample_steps")
    ap.add_argument("--workload", default="text2timbre", choices=["text2timbre", "sharded1024", "modify"],
                    help="text2timbre = BASELINE configs[2] (default, the driver's line); sharded1024 = configs[3]; modify = configs[4]")
    ap.add_argument("--max-batch", type=int, default=0, dest="max_batch", help="modify: largest batch of the sweep (default 4096 on >= 8 GPUs, else 1024)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
        {"text2timbre": run_b200, "sharded1024": run_sharded1024, "modify": run_modify}[args.workload](args)


if __name__ == "__main__":
    main()
