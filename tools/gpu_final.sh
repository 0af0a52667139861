# One gpurun call that produces everything tools/make_profiles.py needs for a round tag:
#   TAG=r1h bash tools/gpu_final.sh
# (bench lines first, each to completion without a profiler; then the two ncu passes)
set -x
G=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $G/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $G/pytest_$TAG.log
timeout 900 python bench.py > $G/bench_$TAG.json 2> $G/bench_$TAG.err
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-e2e --cpu-seconds 10 > $G/bench_${TAG}_cfg3.json 2> $G/bench_${TAG}_cfg3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $G/bench_${TAG}_reference.json 2> $G/bench_${TAG}_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $G/launches_$TAG.csv \
    python bench.py --no-e2e --no-cpu --channels 8 --steps 2 --warmup 1 > $G/ncu_launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_ -c 8 -f -o $G/prof_fused_$TAG \
    python bench.py --no-e2e --no-cpu --channels 16 --steps 1 --warmup 1 > $G/ncu_full_$TAG.log 2>&1
tail -2 $G/pytest_$TAG.log; cat $G/bench_$TAG.json | cut -c1-300
