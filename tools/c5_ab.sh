# A/B timings of the C5 step (cd_blk_kernel) under its build / path switches: tools/c5_ab.sh "VAR=val ..." "VAR=val ..."
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['ms_per_step']))"
done
