for mb in 16 32 64 128; do
  for th in 16; do
    AWRY_B200_CHUNK_MB=$mb AWRY_B200_HOST_THREADS=$th python bench.py --steps 5 --warmup 3 --no-locate --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb $mb threads $th e2e ms', round(d['e2e']['ms_per_step'],2), 'M reads/s', round(d['e2e']['value']/1e6,1))"
  done
done
AWRY_B200_CHUNK_MB=64 AWRY_B200_HOST_THREADS=8 python bench.py --steps 5 --warmup 3 --no-locate --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb 64 threads 8 e2e ms', round(d['e2e']['ms_per_step'],2))"
AWRY_B200_HOST_PACK=0 python bench.py --steps 5 --warmup 3 --no-locate --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ascii e2e ms', round(d['e2e']['ms_per_step'],2))"
