# A/B on one box: round-1 tree (_r1ab) vs the working tree, guard on / off, plus ncu launch lists
G=$PWD/gpurun_out/$TAG; mkdir -p $G
(cd _r1ab && python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > $G/r1_cfg2.json 2> $G/r1_cfg2.err)
python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 --no-guard > $G/noguard_cfg2.json 2> $G/noguard_cfg2.err
python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > $G/guard_cfg2.json 2> $G/guard_cfg2.err
(cd _r1ab && python bench.py --no-e2e --no-cpu --workload cfg3 --steps 3 --warmup 2 > $G/r1_cfg3.json 2> $G/r1_cfg3.err)
python bench.py --no-e2e --no-cpu --workload cfg3 --steps 3 --warmup 2 --no-guard > $G/noguard_cfg3.json 2> $G/noguard_cfg3.err
python bench.py --no-e2e --no-cpu --workload cfg3 --steps 3 --warmup 2 > $G/guard_cfg3.json 2> $G/guard_cfg3.err
python tools/show_bench.py $G/*.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $G/launches_guard.csv python bench.py --no-e2e --no-cpu --channels 8 --steps 2 --warmup 1 > $G/ncu1.log 2>&1
(cd _r1ab && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $G/launches_r1.csv python bench.py --no-e2e --no-cpu --channels 8 --steps 2 --warmup 1 > $G/ncu2.log 2>&1)
