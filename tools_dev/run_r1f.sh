mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
tail -3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt; tail -2 gpurun_out/smoke.log
cat gpurun_out/rc.txt
