python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for f in 1 0; do echo "RN_SVD_FUSED=$f"; RN_SVD_FUSED=$f python tools/svd_time.py 2>&1 | grep "precond=True"; done
python tools/pyprof_dmrg.py 512 20 holstein_dmrg 3 2>&1 | head -12
