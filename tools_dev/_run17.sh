mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_attn_fused.py tests/test_gpu_abi_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -x > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"; tail -n 3 gpurun_out/t_kern.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_headline.py -q -m gpu --tb=short -x > gpurun_out/t_model.log 2>&1; echo "model rc=$?"; tail -n 3 gpurun_out/t_model.log
for rep in 1 2 3; do
echo "new:  $(timeout 300 python tools_dev/step_time.py)"
echo "prev: $(DS_LIB_PATH=$PWD/build/lib_prev.so timeout 300 python tools_dev/step_time.py)"
done 2>&1 | tee gpurun_out/step_ab.log
