# scratch driver for one gpurun call: [parity tests,] bench lines of the default build and of the
# experimental variants under ghost_b200/variants/ (tools/build_variant.sh), per-class times on stderr
mkdir -p gpurun_out/$TAG
if [ -n "$PYTEST" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$TAG/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$TAG/pytest.log; tail -3 gpurun_out/$TAG/pytest.log; fi
export GCWT_CLASS_TIMES=1
for v in default $VARIANTS; do
  if [ $v = default ]; then unset GCWT_LIB; else export GCWT_LIB=$PWD/ghost_b200/variants/libghostcwt_$v.so; fi
  timeout 300 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/$TAG/bench_cfg2_$v.json 2> gpurun_out/$TAG/bench_cfg2_$v.err
  timeout 300 python bench.py --no-e2e --no-cpu --workload cfg3 --steps 2 --warmup 1 > gpurun_out/$TAG/bench_cfg3_$v.json 2> gpurun_out/$TAG/bench_cfg3_$v.err
done
unset GCWT_LIB GCWT_CLASS_TIMES
if [ -n "$FUZZ" ]; then timeout 600 python tools/fuzz_fp32_vs_fp64.py $FUZZ > gpurun_out/$TAG/fuzz.log 2>&1; fi
