timeout 600 python -m pytest tests/test_gpu_zz_training.py -q -m gpu -x --tb=short 2>&1 | grep -v "^$" | tail -15 | cut -c1-400
