for dpt in 3 4; do echo "== depth $dpt"; DINT_B200_SYNC_DEPTH=$dpt timeout 200 python profiles/tools/discover_knobs.py 256 : 2>&1 | tail -4; done
