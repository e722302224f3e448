#!/bin/bash
# End-of-round measurement pass (1 GPU): parity tests, smoke, the full bench line (headline + sub_results), the
# reference arm, the H_eff roofline at three bond dimensions, ncu launch lists of one timed step of the headline
# workload and of one converged Holstein sweep.  ~12 min; the launch lists take most of it.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_pytest.log 2>&1
tail -4 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -2 gpurun_out/r2f_bench.err; python tools/show_bench.py gpurun_out/r2f_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
cut -c1-400 gpurun_out/r2f_bench_ref.json
for M in 256 512 1024; do timeout 300 python tools/hop_roofline.py $M 2>&1 | grep "path=1" ; done > gpurun_out/r2f_hop.log
cat gpurun_out/r2f_hop.log
timeout 1200 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-sub --no-e2e --no-roofline --no-cpu-baseline --profiler-range > gpurun_out/r2f_ncu2.log 2>&1
python tools/launch_summary.py gpurun_out/r2f_bench_launches.csv > gpurun_out/r2f_launch_summary.md
head -24 gpurun_out/r2f_launch_summary.md
# (a launch list of a Holstein sweep takes > 20 min under ncu: ~50 k launches; tools/pyprof_dmrg.py gives the breakdown instead)
timeout 240 python tools/pyprof_dmrg.py 512 20 holstein_dmrg 3 > gpurun_out/r2f_dmrg_pyprof.txt 2>&1; head -24 gpurun_out/r2f_dmrg_pyprof.txt
