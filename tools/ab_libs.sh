# A/B of library builds in one GPU session: bash tools/ab_libs.sh "<bench_ops args>;..." lib1.so lib2.so ...
IFS=';' read -ra SHAPES <<< "$1"; shift
for round in 1 2; do
  for s in "${SHAPES[@]}"; do
    for lib in "$@"; do
      echo -n "$(basename $lib) r$round: "; B200_LIB=$PWD/$lib python tools/bench_ops.py $s 2>&1 | tail -1
    done
  done
done
