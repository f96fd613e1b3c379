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
def cpu_arm(sample_steps: int, unet_evals: int = 2, threads: int | None = None):
    from diffusynth_b200 import weights as W
    from oracle import ds_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    usd, vsd = W.unet_random_state_dict(seed=0, perturb_norm=False), W.vqgan_random_state_dict(seed=1, perturb_norm=False)
    _, dec_plan = W.vqgan_layer_plan(W.VQGAN_DEPLOYED)
    cond, uncond = W.synthetic_conditions(1, 512)
    sch = O.Schedule(1000)
    sch.respace(list(np.linspace(0, 999, unet_evals, dtype=np.int32)))
    draws = W.host_noise(0, 1 + unet_evals, 1)
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
        O.spectrogram_to_waveform(spec[0].numpy().astype(np.float64))
        t4 = time.perf_counter()
    step_s = (t1 - t0) / unet_evals
    tail_s = t4 - t1
    per_timbre = step_s * sample_steps + tail_s
    return dict(value=1.0 / per_timbre, unit="timbres/s", cores=threads, kind="port",
                sample=f"batch 1: {unet_evals} CFG-doubled U-Net steps ({step_s:.3f} s each) + VQ {t2 - t1:.3f} s + decoder {t3 - t2:.3f} s + iSTFT {t4 - t3:.3f} s, "
                       f"extrapolated to {sample_steps} steps; fp32 torch oracle port of the reference",
                s_per_unet_step=step_s, s_tail=tail_s, s_per_timbre=per_timbre)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_arm(args.sample_steps, unet_evals=1)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_arm(args.sample_steps, unet_evals=2)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = dict(metric=METRIC, value=v, unit="timbres/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * wall / max(1, args.steps), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(args, args.gpus),
                cpu_baseline=dict(value=v, unit="timbres/s", cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=v, unit="timbres/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(json.dumps(line))


def workload_config(args, world):
    return dict(workload=f"text-to-timbre full pipeline (BASELINE configs[2]): batch {BATCH}/GPU, {args.sample_steps} DDIM steps, CFG {CFG_SCALE}, "
                         f"latent 4x128x{WIDTH} -> VQ(8192) -> decoder [3,512,{4 * WIDTH}] -> iSTFT {256 * (4 * WIDTH - 1)} samples; deployed U-Net (106.9M) + VQGAN",
                batch_per_gpu=BATCH, global_batch=BATCH * world, sample_steps=args.sample_steps, cfg_scale=CFG_SCALE, sampler="ddim",
                parallelism=f"dp{world} (prompts sharded, final all-gather of waveforms)" if world > 1 else "single GPU",
                l2_policy="per-step working set (multi-GB of activations) exceeds the 126 MB L2; no explicit flush")


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def conv_flops(a) -> float:
    """Algorithmic FLOPs (2/MAC, real channels, no padding) of one ds_conv_gemm argument block."""
    k = a.ntaps * (a.C0 + a.C1)
    return 2.0 * a.N * a.groups * a.H * a.W * a.Cout * k


def profile_unet_eval(plan):
    """One eager (un-graphed) U-Net evaluation with CUDA events around every launch -> ms per kernel family and the
    algorithmic FLOPs of the tcgen05 conv launches."""
    from diffusynth_b200._lib import ConvGemmArgs
    evs = []
    torch.cuda.synchronize()
    for name, fn in plan.ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((name, e0, e1))
    torch.cuda.synchronize()
    fam = {}
    for name, e0, e1 in evs:
        if False:
            pass
        elif name.endswith(("net.1", "net.4", "final_conv.1")):
            key = "conv_gemm(3x3)"
        elif name.endswith(("res_conv", "to_qkv", "to_out")):
            key = "conv_gemm(1x1)"
        elif name.endswith(("ds_conv",)):
            key = "dwconv7"
        elif name.endswith(("ctx", "fin")):
            key = "attn_core"
        elif name.endswith("gn_res"):
            key = "gn_apply_residual"
        elif name == "init_conv":
            key = "conv_gemm(1x1)"
        elif name == "init_im2col":
            key = "stem_im2col"
        elif name.startswith("time_"):
            key = "time_mlp"
        else:
            key = "conv_gemm(4x4s2/T)"
        fam[key] = fam.get(key, 0.0) + e0.elapsed_time(e1)
    flops = sum(conv_flops(k) for k in plan.keep if isinstance(k, ConvGemmArgs))
    launches = sum(1 for k in plan.keep if isinstance(k, ConvGemmArgs))
    if os.environ.get("DS_DUMP_OPS"):
        convs = [k for k in plan.keep if isinstance(k, ConvGemmArgs)]
        ci, rows = 0, []
        for name, e0, e1 in evs:
            ms = e0.elapsed_time(e1)
            row = dict(op=name, ms=round(ms, 4))
            is_conv = name == "init_conv" or name.endswith(("net.1", "net.4", "final_conv.1", "res_conv", "to_qkv", "to_out")) or name.split(".")[-1] in ("2", "4") and name.count(".") == 2
            if is_conv and ci < len(convs):
                a = convs[ci]; ci += 1
                row.update(tflops=round(conv_flops(a) / (ms / 1e3) / 1e12, 1), H=a.H, W=a.W, Cin=a.C0 + a.C1, Cout=a.Cout, taps=a.ntaps, BN=a.BN, BK=a.BK)
            rows.append(row)
        with open(os.environ["DS_DUMP_OPS"], "w") as f:
            json.dump(rows, f, indent=0)
    return fam, flops, launches


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

    def step_device():
        out = pipe.generate(cond_dev, uncond_dev, steps=args.sample_steps, cfg_scale=CFG_SCALE, width=WIDTH, seed=None)
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
    for _ in range(max(3, args.warmup)):
        step_device()
    with ClockSampler(local) as clk:
        ms_dev = timed(step_device, args.steps)
    clocks = clk.summary()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    value = total * args.steps / (ms_dev / 1e3)
    e2e_value = total * args.steps / (ms_e2e / 1e3)
    line = None
    if rank == 0:
        pk = peaks()
        sampler = pipe._samplers[(BATCH, args.sample_steps)]
        loop = next(iter(sampler._graphs.values()))
        plan = loop.plan
        fam, flops, conv_launches = profile_unet_eval(plan)
        fam2, _, _ = profile_unet_eval(plan)
        fam = {k: min(v, fam2[k]) for k, v in fam.items()}
        conv_ms = sum(v for k, v in fam.items() if k.startswith("conv_gemm"))
        unet_ms = sum(fam.values())
        achieved = flops / (conv_ms / 1e3) / 1e12
        # time of the sampling graph alone (U-Net step ms)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            loop.launch()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / 3 / loop.n_iter
        algo_step_tflop = 2 * BATCH * 136.70e9 / 1e12
        dec_plan = next(iter(pipe.vqgan._decoder._stack._plans.values()))
        launches = loop.launches + len(plan.cond_ops) + 1 + len(dec_plan.ops) + 2
        cpu = cpu_arm(args.sample_steps, unet_evals=2) if world == 1 else None
        line = dict(metric=METRIC, value=value, unit="timbres/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp16", data="synthetic",
                    config=workload_config(args, world), clocks=clocks,
                    e2e=dict(value=e2e_value, unit="timbres/s", h2d_bytes_per_step=int(cond_host.numel() * 4),
                             d2h_bytes_per_step=int(wave_host.numel() * 4), ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(launches * args.steps),
                    unet_step_ms=step_ms,
                    unet_step_roofline=dict(algorithmic_tflop=algo_step_tflop, achieved_tflops=algo_step_tflop / (step_ms / 1e3),
                                            peak_tflops=pk["tf_sustained"], frac=algo_step_tflop / (step_ms / 1e3) / pk["tf_sustained"]),
                    roofline=dict(kernel="conv_gemm_kernel (tcgen05 implicit-GEMM conv; all dense convs of the U-Net)", bound="tensor",
                                  achieved=achieved, peak=pk["tf_sustained"], unit="TFLOP/s", frac=achieved / pk["tf_sustained"], traffic=conv_traffic(),
                                  traffic_note="ncu dram__bytes_read+write per conv_gemm launch, averaged over the 98 launches of one U-Net evaluation "
                                               "(profiles/r01_unet_eval_ncu_summary.md)",
                                  peak_source=pk["source"] + " bf16_tflops_sustained", launches_per_unet_eval=conv_launches,
                                  algorithmic_gflop_per_unet_eval=flops / 1e9, ms_per_unet_eval=conv_ms),
                    kernel_ms_per_unet_eval={k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
                    unet_eval_ms_eager=unet_ms)
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_conv_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["avg_dram_bytes_per_conv_launch"])
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sample-steps", type=int, default=20, dest="sample_steps")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
        run_b200(args)


if __name__ == "__main__":
    main()
