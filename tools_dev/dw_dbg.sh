timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -s -k "stft or griffin or glue or pipeline or modification or note or round_trip" 2>&1 | grep -E "passed|failed|Error|assert|griffin|spectral|pipeline:|modification:|note" | tail -12
python - <<'PY'
import torch, time
from diffusynth_b200 import codec
spec = torch.randn(64, 3, 512, 256, device="cuda"); spec[:, 0].abs_()
w = codec.spectrogram_to_waveform(spec); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): w = codec.spectrogram_to_waveform(spec)
e1.record(); torch.cuda.synchronize(); print("decode+istft batch 64: %.3f ms" % (e0.elapsed_time(e1) / 20))
e0.record()
for _ in range(20): s2 = codec.waveform_to_spectrogram(w)
e1.record(); torch.cuda.synchronize(); print("stft+encode batch 64: %.3f ms" % (e0.elapsed_time(e1) / 20))
PY
