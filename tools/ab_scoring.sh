for rep in 1 2; do
for a in "10000 50000 128 100 10" "10000 1000000 256 100 4" "10000 6250 128 100 10"; do
  echo "new: $(python tools/score_bench.py $a 2>&1 | tail -1)"
  echo "old: $(SERT_B200_LIB=$PWD/sert_b200/libsert_prev.so python tools/score_bench.py $a 2>&1 | tail -1)"
done
done
