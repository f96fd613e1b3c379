"""The non-U-Net tail of the pipeline once, eagerly (for ncu): one sampling step (fused CFG + DDIM update), VQ, VQGAN decoder,
STFT+ decode + iSTFT, and the encoder side (STFT encode + VQGAN encoder) at batch 64."""
import os, sys
os.environ["DS_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import TextToTimbre, weights as W
B = int(os.environ.get("B", "64"))
pipe = TextToTimbre.random_init(device="cuda", seed=0)
cond, uncond = W.synthetic_conditions(B, 512)
out = pipe.generate(cond.cuda(), uncond.cuda(), steps=1, cfg_scale=6, seed=0)
lat = pipe.encode_audio(out.waveforms[:B])
torch.cuda.synchronize()
print("ok", tuple(out.waveforms.shape), tuple(lat.shape), float(out.waveforms.abs().mean()))
