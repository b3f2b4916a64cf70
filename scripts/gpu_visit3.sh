#!/bin/bash
# GPU visit: sanitizer on small proofs (both diagonal forms), parity suite, smoke, bench
mkdir -p gpurun_out
for k in 1 0; do
MP_SMALL_DECK_MAX=0 MP_DIAG_KARATSUBA=$k timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python scripts/repro.py 5 3 2>&1 | grep -v "Host Frame\|^=========         in " | tail -4
done
MP_SMALL_DECK_MAX=0 MP_DIAG_KARATSUBA=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python scripts/repro.py 4 4 2>&1 | grep -v "Host Frame\|^=========         in " | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 600 gpurun_out/bench.json; echo; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['split']); print(d['pipelined']); print(d['roofline'])
PY
tail -5 gpurun_out/bench.err
