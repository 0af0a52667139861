# time experimental builds (ghost_b200/variants/libghostcwt_<v>.so) against the default: VARIANTS="a b" bash tools/gpu_variants.sh
G=gpurun_out/$TAG; mkdir -p $G
for v in default $VARIANTS; do
  if [ $v = default ]; then unset GCWT_LIB; else export GCWT_LIB=$PWD/ghost_b200/variants/libghostcwt_$v.so; fi
  python bench.py --no-e2e --no-cpu --no-guard --steps 10 --warmup 3 $EXTRA > $G/cfg2_$v.json 2> $G/cfg2_$v.err
done
unset GCWT_LIB
python tools/show_bench.py $G/cfg2_*.json
