# e2e (C-ABI call, pinned host buffers) vs the share of the query bytes packed on the host
for share in auto 1 0.8 0.7 0.6 0.5 0; do
  if [ "$share" = auto ]; then unset AWRY_B200_PACK_SHARE; else export AWRY_B200_PACK_SHARE=$share; fi
  python bench.py --steps 5 --warmup 3 --no-locate --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pack share $share: e2e ms', round(d['e2e']['ms_per_step'],2), 'M reads/s', round(d['e2e']['value']/1e6,1), 'h2d MB', d['e2e']['h2d_bytes_per_step']//1000000, ' device-resident ms', round(d['ms_per_step'],2))"
done
