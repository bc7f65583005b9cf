# cfg2 end to end from pinned ASCII (scripts/count_e2e_trace.py: 12 calls, median of the last 8) against the share of
# the bytes the host packs (AWRY_B200_PACK_SHARE) and the chunk size (AWRY_B200_CHUNK_MB)
for chunk in 128 64; do
for share in auto 1 0.9 0.8 0.7 0.6; do
  if [ "$share" = auto ]; then unset AWRY_B200_PACK_SHARE; else export AWRY_B200_PACK_SHARE=$share; fi
  AWRY_B200_CHUNK_MB=$chunk python scripts/count_e2e_trace.py 2>&1 | python -c "
import sys,re,statistics
v=[float(m.group(1)) for m in (re.search(r'call \d+: ([0-9.]+) ms',l) for l in sys.stdin) if m][4:]
print('chunk $chunk MB, pack share $share: median', round(statistics.median(v),2), 'ms, min', round(min(v),2))"
done; done
