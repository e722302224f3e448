( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload holstein_dmrg --no-cpu-baseline --no-e2e --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('holstein value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline'] and d['roofline']['frac'])"
timeout 200 python tools/pyprof_dmrg.py 512 20 holstein_dmrg 3 > gpurun_out/r2f_dmrg_pyprof.txt 2>&1; head -3 gpurun_out/r2f_dmrg_pyprof.txt | cut -c1-200
