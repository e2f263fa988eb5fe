for v in "" cc; do
  if [ -z "$v" ]; then unset OBIA_B200_LIB; n=main; else export OBIA_B200_LIB=/root/repo/obia_b200/_lib/variants/lib_$v.so; n=$v; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-tiled --no-alt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$n', round(d['ms_per_step'],2), d['split_ms'])"
done
