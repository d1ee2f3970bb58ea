set -x
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q -k "index or interpolative or block128" 2>&1 | tail -3
python profiles/tools/small_configs.py > gpurun_out/small_configs.json 2> gpurun_out/small_configs.err; tail -2 gpurun_out/small_configs.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/small_configs.json'))
for k,v in d['config2_index_scan'].items(): print(k, round(v['us_per_pass'],1),'us', round(v['G_ints_per_s'],1),'G ints/s')
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/index_launches.csv python profiles/tools/index_profile.py > gpurun_out/index_profile.log 2>&1
grep -v "^==" gpurun_out/index_launches.csv > gpurun_out/index_launches_clean.csv
python - <<'PY'
import csv,collections
agg=collections.OrderedDict()
for r in csv.DictReader(open('gpurun_out/index_launches_clean.csv')):
    a=agg.setdefault(r['Kernel Name'][:60],[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])
for k,a in agg.items(): print('%10.0f ns avg x%d %s'%(a[1]/a[0],a[0],k))
PY
