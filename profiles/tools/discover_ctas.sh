for n in 3 4; do echo "== count CTAs per SM $n"; DINT_B200_COUNT_CTAS=$n timeout 200 python profiles/tools/discover_knobs.py 256 : 2>&1 | tail -4; done
