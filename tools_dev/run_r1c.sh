# round-1 late session: new features (odd widths, image products, note synthesis) + dwconv HFMA2 A/B
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_kernels.py -q -m gpu --tb=short > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?" >> gpurun_out/rc.txt
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
tail -15 gpurun_out/t_kern.log; grep -E "rel-L2|latent|passed|failed|Error|error" gpurun_out/t_model.log | tail -30
timeout 300 python tools_dev/ab_dwconv.py > gpurun_out/ab_dw_f32.log 2>&1; echo "ab0 rc=$?" >> gpurun_out/rc.txt
DS_DWCONV_HACC=1 timeout 300 python tools_dev/ab_dwconv.py > gpurun_out/ab_dw_hacc.log 2>&1; echo "ab1 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/ab_dw_f32.log gpurun_out/ab_dw_hacc.log
DS_DWCONV_HACC=1 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -s -k "unet_forward_parity or graph_sampling_loop or text_to_timbre" > gpurun_out/t_model_hacc.log 2>&1; echo "model_hacc rc=$?" >> gpurun_out/rc.txt
grep -E "rel-L2|latent|passed|failed" gpurun_out/t_model_hacc.log | grep -v "^    " | tail -12
DS_DWCONV_HACC=1 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_hacc.json 2> gpurun_out/bench_hacc.err; echo "bench_hacc rc=$?" >> gpurun_out/rc.txt
python -c "
import json; d=json.loads(open('gpurun_out/bench_hacc.json').read().strip().splitlines()[-1]); print('HACC bench', d['value'], d['unet_step_ms'], d['kernel_ms_per_unet_eval'])"
cat gpurun_out/rc.txt
