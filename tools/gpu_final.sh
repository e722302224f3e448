#!/bin/bash
# End-of-round measurement pass (1 GPU): parity tests, smoke, bench line (default workload + Holstein
# DMRG M=512), H_eff roofline at three bond dimensions, ncu --set full of the GEMM at M=1024 and of the
# block-Jacobi kernels, ncu launch list of one timed bench step.
# Budget: everything up to the last step takes ~4 min; the launch list (8.5 k launches under ncu) takes
# ~12 min on its own -- run it as a separate gpurun call when fewer than 20 GPU-minutes are left.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
tail -4 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
cat gpurun_out/f_bench.json; tail -2 gpurun_out/f_bench.err
timeout 600 python bench.py --workload holstein_dmrg --bond 512 --steps 2 --warmup 1 --no-e2e --cpu-budget 20 > gpurun_out/f_bench_dmrg512.json 2> gpurun_out/f_bench_dmrg512.err
cat gpurun_out/f_bench_dmrg512.json; tail -2 gpurun_out/f_bench_dmrg512.err
for M in 256 512 1024; do timeout 300 python tools/hop_roofline.py $M 2>&1 | grep "path=1" ; done > gpurun_out/f_hop.log
cat gpurun_out/f_hop.log
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:jb_ -s 96 -c 3 --profile-from-start off -o gpurun_out/f_jb python tools/svd_one.py > gpurun_out/f_ncu3.log 2>&1
tail -1 gpurun_out/f_ncu3.log
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ozaki_gemm_kernel -s 2 -c 2 -o gpurun_out/f_gemm1024 python tools/hop_roofline.py 1024 > gpurun_out/f_ncu1.log 2>&1
tail -1 gpurun_out/f_ncu1.log
timeout 1500 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-roofline --no-cpu-baseline --profiler-range > gpurun_out/f_ncu2.log 2>&1
python tools/launch_summary.py gpurun_out/f_bench_launches.csv > gpurun_out/f_launch_summary.md
head -30 gpurun_out/f_launch_summary.md
