mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
tail -3 gpurun_out/t_all.log
DS_DUMP_OPS=gpurun_out/ops.json timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
tail -c 700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat gpurun_out/rc.txt
