# chunk size x pool size for the wire way home of bsgpu_call_sites (BSGPU_WIRE: 0 full records, 1 adaptive, 2 wire only)
set -u
export E2E_ONLY_SITES=1
for c in 65536 131072 262144; do
  for t in 8 10 12; do
    BSGPU_SITES_CHUNK=$c BSGPU_WIRE=2 BSGPU_EXPAND_THREADS=$t timeout 300 python profiles/e2e_only.py 1.6e7 2>&1 | tail -1
  done
done
for t in 6 8 10; do
BSGPU_WIRE=1 BSGPU_EXPAND_THREADS=$t timeout 300 python profiles/e2e_only.py 1.6e7 2>&1 | tail -1
done
BSGPU_WIRE=0 timeout 300 python profiles/e2e_only.py 1.6e7 2>&1 | tail -1
