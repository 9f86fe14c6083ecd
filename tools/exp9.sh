cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/bench_configs.py c1 c3_table c2_expand c5_n2 c5_n3 c5_n5 c5_n6 c5_n7 c5_density_w16 c5_density_w8 c5_density_w4 c5_density_w2 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    for k, v in d.items(): print('%-20s frac=%.3f value=%.3g %s' % (k, v['frac'], v['value'], v.get('ms') or v.get('us_per_step')))
"
