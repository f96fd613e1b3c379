export REPS=1
timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_full.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_toout2 -f python tools_dev/unet_once.py > gpurun_out/ncu_toout.log 2>&1
tail -2 gpurun_out/ncu_toout.log
timeout 200 python tools_dev/ab_attn.py
