mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
tail -4 gpurun_out/t_all.log
DS_DUMP_OPS=gpurun_out/ops.json timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
export REPS=1
timeout 600 python tools_dev/unet_once.py > gpurun_out/plain_traffic.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/traffic.csv \
    python tools_dev/unet_once.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncutraffic rc=$?" >> gpurun_out/rc.txt
tail -2 gpurun_out/ncu_traffic.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt; tail -2 gpurun_out/smoke.log
cat gpurun_out/rc.txt
