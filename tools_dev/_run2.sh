mkdir -p gpurun_out
for rep in 1 2; do
DS_PDL=1 timeout 300 python tools_dev/step_time.py
DS_PDL=0 timeout 300 python tools_dev/step_time.py
done 2>&1 | tee gpurun_out/pdl_ab.log
timeout 900 python -m pytest tests/test_gpu_attn_fused.py tests/test_gpu_conv_gemm.py tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"; tail -5 gpurun_out/t_kern.log
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_sampler_surfaces.py -q -m gpu --tb=short -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -5 gpurun_out/t_model.log
