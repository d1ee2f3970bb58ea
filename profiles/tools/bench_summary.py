import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r=d['roofline']
    print(f, "value %.1f ms/step %.2f docs %.2f ms (%.3f) freqs %.2f ms (%.3f) e2e %s" % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['freqs_kernel']['avg_launch_ms'], r['freqs_kernel']['achieved']/r['peak'], d.get('e2e',{}).get('value')))
