timeout 900 python -m pytest tests/test_ablations_gpu.py -x -q 2>&1 | tail -30
