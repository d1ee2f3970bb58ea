for K in 1 8 32; do echo "== K=$K"; timeout 200 python profiles/tools/discover_knobs.py $K : 8192:4096 12288:6144 16384:8192 2>&1 | grep '"rep": 1' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['part'], r['long'] or 'dflt', r['cut'] or 'dflt', r['ms'], r['phases'], r['chains'], r['skimmed'], r['retries'])"; done
