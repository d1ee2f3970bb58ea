set -x
cd /root/repo
DINT_B200_STRIP_VARIANT=4 timeout 300 python -m pytest tests -m gpu -x -q -k "vroom_raw or fused_docids or random_streams or unaligned or edge_case" 2>&1 | tail -3
DINT_B200_STRIP_VARIANT=7 timeout 300 python -m pytest tests -m gpu -x -q -k "vroom_raw or fused_docids or random_streams" 2>&1 | tail -3
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash profiles/tools/sweep.sh DINT_B200_STRIP_VARIANT=0 DINT_B200_STRIP_VARIANT=4 DINT_B200_STRIP_VARIANT=3 DINT_B200_STRIP_VARIANT=7
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/index_launches.csv python profiles/tools/index_profile.py > gpurun_out/index_profile.log 2>&1
grep -v "^==" gpurun_out/index_launches.csv | python -c "
import csv,sys,collections
agg=collections.OrderedDict()
for r in csv.DictReader(sys.stdin):
    a=agg.setdefault(r['Kernel Name'][:60],[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])
for k,a in agg.items(): print('%10.0f ns avg x%d %s'%(a[1]/a[0],a[0],k))
"
