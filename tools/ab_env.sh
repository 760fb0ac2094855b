# A/B of an environment switch in one GPU session: bash tools/ab_env.sh VAR "v1 v2 .." "<bench_ops args>;..."
VAR=$1; VALS=$2; IFS=';' read -ra SHAPES <<< "$3"
for s in "${SHAPES[@]}"; do
  for v in $VALS; do
    echo -n "$VAR=$v: "; env $VAR=$v python tools/bench_ops.py $s 2>&1 | tail -1
  done
done
