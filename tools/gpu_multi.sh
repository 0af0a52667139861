# N-GPU checks in one gpurun --gpus N call: NG=2 bash tools/gpu_multi.sh
G=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --no-e2e --no-cpu > $G/bench_multi_cfg2_n$NG.json 2> $G/bench_multi_cfg2_n$NG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --workload cfg4 --steps 2 --warmup 1 --no-e2e --no-cpu > $G/bench_multi_cfg4_n$NG.json 2> $G/bench_multi_cfg4_n$NG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/check_time_shard.py > $G/check_time_shard_n$NG.log 2>&1
tail -3 $G/check_time_shard_n$NG.log
