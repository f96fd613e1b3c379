mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "all rc=$?" > gpurun_out/rc.txt; tail -5 gpurun_out/t_all.log
