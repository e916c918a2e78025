bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/variants3.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/gpu_bvh_perf.py 32 2>&1 | tail -12
