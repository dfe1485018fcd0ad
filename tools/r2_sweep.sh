timeout 1200 python -m pytest tests/test_main_gpu.py tests/test_engine_gpu.py -x -q -s > gpurun_out/main_test.log 2>&1
grep -n "main.py yelp shape" gpurun_out/main_test.log; tail -25 gpurun_out/main_test.log
