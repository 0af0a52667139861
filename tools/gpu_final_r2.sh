# One gpurun call that produces everything tools/make_profiles_r2.py needs:  TAG=r2a bash tools/gpu_final_r2.sh
# (bench lines first, each to completion without a profiler; then the ncu passes)
set -x
G=gpurun_out/$TAG; mkdir -p $G
timeout 1200 python -m pytest tests -m gpu -q > $G/pytest.log 2>&1; echo "pytest rc=$?" >> $G/pytest.log
timeout 900 python bench.py > $G/bench_cfg2.json 2> $G/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $G/bench_reference.json 2> $G/bench_reference.err
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --cpu-seconds 10 > $G/bench_cfg3_shard.json 2> $G/bench_cfg3_shard.err
timeout 600 python bench.py --dtype f64 --steps 5 --warmup 3 --cpu-seconds 10 > $G/bench_f64.json 2> $G/bench_f64.err
timeout 600 python bench.py --no-guard --no-e2e --no-cpu > $G/bench_cfg2_noguard.json 2> $G/bench_cfg2_noguard.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $G/launches_cfg2_8ch.csv \
    python bench.py --no-e2e --no-cpu --channels 8 --steps 2 --warmup 1 > $G/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fused_|pyramid_kernel|mean_partial|guard_eval" -c 24 -f -o $G/prof_fused \
    python bench.py --no-e2e --no-cpu --channels 16 --steps 1 --warmup 1 > $G/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fs_|generic_response" -c 6 -f -o $G/prof_f64 \
    python bench.py --dtype f64 --no-e2e --no-cpu --steps 1 --warmup 1 > $G/ncu_f64.log 2>&1
tail -2 $G/pytest.log; cut -c1-300 $G/bench_cfg2.json
