mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_kernels.py -q -m gpu --tb=short > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?" >> gpurun_out/rc.txt
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
tail -6 gpurun_out/t_kern.log; grep -E "rel-L2|latent|passed|failed|Error|error|griffin|spectral" gpurun_out/t_model.log gpurun_out/t_kern.log | grep -v "^gpurun_out/t_model.log:    " | tail -32
DS_DUMP_OPS=gpurun_out/ops.json timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['unet_step_ms'], d['kernel_ms_per_unet_eval'], d['clocks'])"
tail -3 gpurun_out/bench.err
cat gpurun_out/rc.txt
