cd $GRAFT_REPO_ROOT
for t in ro256x3 ro256x2 ro512; do echo "== $t"; MAPF_B200_LIB=gym_mapf_b200/csrc/libmapf_b200_$t.so python tools/bench_configs.py c2_rollout c2_rollout_random 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line.strip()[:200]); continue
    for k, v in d.items(): print('%-20s frac=%.3f value=%.3g %s' % (k, v['frac'], v['value'], v.get('ms') or v.get('us_per_step')))
"; done
