for c in "" "1,2,4,6,6,6" "1,1,3,5,7,8" "1,2,3,5,7,7" "2,3,5,7,8" "1,1,2,3,4,6,8"; do
  echo "== chunks: ${c:-default 1,1,2,4,5,6,6}"
  SPG_UPLOAD_CHUNKS=$c python bench.py --no-cpu --no-aux --no-verify 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 3), round(d['e2e']['ms_per_step'], 3), d['proof_sha256'][:12])"
done
