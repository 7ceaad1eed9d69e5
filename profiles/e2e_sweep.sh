#!/bin/bash
# end-to-end pipeline experiments: early D2H panels / panelled uploads / pipeline depth / one shared context (bench.py e2e leg)
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extra 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f  e2e %.1f evals/s (%.2f ms)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
"; }
for cfg in "${@:-RDB_BENCH_DEPTH=4}"; do run $cfg; done
