# usage: bash tools_dev/run_multi_gpu.sh N [workloads...]   (on the GPU box, through gpurun --gpus N): one JSON line per workload into gpurun_out/
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for w in "${@:-text2timbre sharded1024 modify}"; do
  for wl in $w; do
    extra=""; [ "$wl" = "text2timbre" ] && extra="--warmup 3"
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --workload $wl --steps 2 $extra \
      > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
    echo "$wl n=$N rc=$?"; head -c 400 gpurun_out/bench_${wl}_n$N.json; echo; tail -n 2 gpurun_out/bench_${wl}_n$N.err | cut -c1-300
  done
done
