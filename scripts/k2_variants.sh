for v in "" 11 00; do
  if [ -z "$v" ]; then unset OBIA_B200_LIB; n=main; else export OBIA_B200_LIB=/root/repo/obia_b200/_lib/variants/lib_$v.so; n=$v; fi
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-tiled 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$n', round(d['ms_per_step'],2), round(r['avg_launch_ms'],3), round(d['also']['assign_kernel_avg_ms'],3))"
done
