for rep in 1 2; do
  for lib in sert_b200/libsert_b200.so sert_b200/libsert_prev.so; do
    echo "$(basename $lib) f32:  $(SERT_B200_LIB=$PWD/$lib python tools/vs_step_bench.py 400 1 2>&1 | tail -1)"
    echo "$(basename $lib) bf16: $(SERT_BENCH_BF16_STATE=1 SERT_B200_LIB=$PWD/$lib python tools/vs_step_bench.py 400 1 2>&1 | tail -1)"
  done
done
