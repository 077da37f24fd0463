timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "at_size" 2>&1 | tail -30 | cut -c1-250
