# usage: bash tools_dev/run_gpu_round.sh [tests|tests_new|bench|ncu|ncutraffic|ncufull|census] ...   (run on the GPU box through gpurun)
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
for what in "$@"; do
case $what in
tests)
  timeout 900 python -m pytest tests/test_gpu_attn_fused.py tests/test_gpu_conv_gemm.py tests/test_gpu_kernels.py -q -s -m gpu --tb=short > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?" >> gpurun_out/rc.txt
  timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_sampler_surfaces.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
  grep -E "fused vs|passed|failed|Error|error" gpurun_out/t_kern.log | tail -20; tail -12 gpurun_out/t_model.log ;;
headline)
  timeout 1500 python -m pytest tests/test_gpu_headline.py -q -m gpu --tb=short -s > gpurun_out/t_headline.log 2>&1; echo "headline rc=$?" >> gpurun_out/rc.txt
  grep -v "^    step" gpurun_out/t_headline.log | tail -40 ;;
abi)
  timeout 900 python -m pytest tests/test_gpu_abi_engine.py -q -m gpu --tb=short -s > gpurun_out/t_abi.log 2>&1; echo "abi rc=$?" >> gpurun_out/rc.txt
  tail -15 gpurun_out/t_abi.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
  tail -3 gpurun_out/smoke.log ;;
bench)
  DS_DUMP_OPS=gpurun_out/ops.json timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
  tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
benchref)
  timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "benchref rc=$?" >> gpurun_out/rc.txt
  tail -c 1500 gpurun_out/bench_ref.json ;;
ncu)
  timeout 600 python bench.py --steps 1 --warmup 1 --sample-steps 2 > gpurun_out/plain.log 2>&1 && \
  DS_NO_GRAPH=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --sample-steps 2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/rc.txt
  tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log ;;
ncutraffic)
  export REPS=1
  timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_traffic.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/traffic.csv \
      python tools_dev/unet_once.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncutraffic rc=$?" >> gpurun_out/rc.txt
  tail -2 gpurun_out/ncu_traffic.log ;;
ncuattn)
  export REPS=1
  timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_full.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_qkv_ctx -s 0 -c 1 -o gpurun_out/prof_attnqkv -f \
      python tools_dev/unet_once.py > gpurun_out/ncu_attn.log 2>&1; echo "ncuattn rc=$?" >> gpurun_out/rc.txt
  tail -3 gpurun_out/ncu_attn.log ;;
ncufull)
  export REPS=1
  timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_full.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 0 -c 3 -o gpurun_out/prof_conv -f \
      python tools_dev/unet_once.py > gpurun_out/ncu_full.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwconv7 -s 0 -c 1 -o gpurun_out/prof_dw -f \
      python tools_dev/unet_once.py >> gpurun_out/ncu_full.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ctx -s 0 -c 1 -o gpurun_out/prof_attn -f \
      python tools_dev/unet_once.py >> gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?" >> gpurun_out/rc.txt
  tail -3 gpurun_out/ncu_full.log ;;
esac
done
cat gpurun_out/rc.txt
