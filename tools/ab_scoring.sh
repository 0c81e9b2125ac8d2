# A/B/C of libsert_b200 builds in one GPU session: tools/ab_scoring.sh [lib ...] (default: current, libsert_prev.so)
LIBS="${@:-sert_b200/libsert_b200.so sert_b200/libsert_prev.so}"
for rep in 1 2; do
for a in "10000 50000 128 100 10" "10000 1000000 256 100 4" "10000 6250 128 100 10"; do
  for lib in $LIBS; do
    echo "$(basename $lib): $(SERT_B200_LIB=$PWD/$lib python tools/score_bench.py $a 2>&1 | tail -1)"
  done
done
done
