for d in 0 1 2 3 4 7; do echo "debug=$d"; PDM_CONV_DEBUG=$d python tools/bench_conv.py --no-torch 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    if 'layer' in r: print('  %-20s %.3f ms' % (r['layer'], r['ms']))
"; done
