for v in "" b448 b224; do for k in 8 16 32; do
  if [ -z "$v" ]; then lib=""; else lib="NPS_B200_LIB=$PWD/nuclear-sim_b200/_lib/libnps_b200_$v.so"; fi
  echo -n "variant=${v:-b64m7} substeps=$k  "
  env $lib python bench.py --steps 12 --warmup 3 --no-cpu-baseline --substeps $k 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e  e2e %.3e  ms/launch %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done; done
