timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "persistent_cg" 2>&1 | tail -12 | cut -c1-300
