set -x
mkdir -p gpurun_out/b1
python -m pytest tests -m gpu -x -q > gpurun_out/b1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b1/pytest.log
for v in default t10 c3; do
  if [ $v = default ]; then unset GCWT_LIB; else export GCWT_LIB=$PWD/ghost_b200/variants/libghostcwt_$v.so; fi
  python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/b1/bench_cfg2_$v.json 2> gpurun_out/b1/bench_cfg2_$v.err
  python bench.py --no-e2e --no-cpu --workload cfg3 --steps 2 --warmup 1 > gpurun_out/b1/bench_cfg3_$v.json 2> gpurun_out/b1/bench_cfg3_$v.err
done
unset GCWT_LIB
timeout 600 python tools/fuzz_fp32_vs_fp64.py 40 1 > gpurun_out/b1/fuzz.log 2>&1
tail -3 gpurun_out/b1/pytest.log; tail -3 gpurun_out/b1/fuzz.log
for f in gpurun_out/b1/bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); r=d['roofline']
print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_step'],3),round(v['frac'],3)) for k,v in r['families'].items()}, r['whole_step']['frac'])
"; done
