"""U-Net step time inside a CUDA graph (5 evaluations per graph, batch 128 guidance-doubled) + eps checksum; DS_PDL / DS_LIB_PATH select A/B builds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ConditionedUnet, weights as W
B = int(os.environ.get("B", "64"))
net = ConditionedUnet(**{k: v for k, v in W.UNET_DEPLOYED.items() if k not in ("out_dim", "time_dim")}, device="cuda",
                      batch_invariant=os.environ.get("WIDE") == "1")      # WIDE=1: always the widest N tile (A/B of the small-job narrowing)
net.load_state_dict(W.unet_random_state_dict(seed=0, perturb_norm=False))
pl = net.plan(2 * B, 128, 64, x_batch_mod=B, uniform_time=True)
torch.manual_seed(0)
pl.x.normal_(); pl.cond.normal_(); pl.t.fill_(500)
pl.run_cond()
for _ in range(2):
    pl.run()
torch.cuda.synchronize()
ref = pl.eps.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(5):
            pl.run()
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    g.replay()
e1.record()
torch.cuda.synchronize()
same = bool(torch.equal(ref, pl.eps))
print(f"B={B} unet step {e0.elapsed_time(e1) / 20:.3f} ms  (graph of 5, 4 replays)  eps mean|.| {float(pl.eps.abs().mean()):.6f} identical_to_eager={same}")
