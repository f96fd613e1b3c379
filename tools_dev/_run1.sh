mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn_fused.py -q -s -m gpu --tb=short > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?" 
tail -8 gpurun_out/t_attn.log
for rep in 1 2; do
echo "== WIP"; timeout 200 python tools_dev/ab_attn.py
echo "== commit"; DS_LIB_PATH=$PWD/build/lib_commit.so timeout 200 python tools_dev/ab_attn.py
done > gpurun_out/ab_attn2.log 2>&1
cat gpurun_out/ab_attn2.log
export REPS=1
timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_toout -f python tools_dev/unet_once.py > gpurun_out/ncu_toout.log 2>&1
tail -3 gpurun_out/ncu_toout.log
