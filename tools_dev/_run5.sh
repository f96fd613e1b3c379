mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu --tb=short -s > gpurun_out/t_comm.log 2>&1; echo "comm rc=$?"; tail -5 gpurun_out/t_comm.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload sharded1024 --steps 2 > gpurun_out/bench_sharded_n2.json 2> gpurun_out/bench_sharded_n2.err; echo "sharded rc=$?"; tail -c 1500 gpurun_out/bench_sharded_n2.json; tail -3 gpurun_out/bench_sharded_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; head -c 600 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
