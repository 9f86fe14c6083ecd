cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for B in 1048576 8388608; do
 TIME_GRAPH=1 TIME_B=$B timeout 120 python tools/time_step.py ship 2>&1 | tail -1
 TIME_STREAMS=2 TIME_GRAPH=1 TIME_B=$B timeout 120 python tools/time_step.py ship_2pools 2>&1 | tail -1
done
python tools/bench_configs.py c2_rollout c4_step c2_expand c3_table c5_n2 c5_n8 c5_n10 c5_density_w2 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.strip()[:200]); continue
    for k, v in d.items(): print('%-20s frac=%.3f value=%.3g %s' % (k, v['frac'], v['value'], v.get('ms') or v.get('us_per_step')))
"
