mkdir -p gpurun_out
for rep in 1 2 3; do
echo "new:  $(timeout 300 python tools_dev/step_time.py)"
echo "prev: $(DS_LIB_PATH=$PWD/build/lib_prev.so timeout 300 python tools_dev/step_time.py)"
done 2>&1 | tee gpurun_out/step_ab.log
