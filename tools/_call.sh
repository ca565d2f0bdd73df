timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_online.py -x -q > gpurun_out/r2_pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_train.log; tail -6 gpurun_out/r2_pytest_train.log
timeout 300 python tools/bench_train.py > gpurun_out/r2_train_bench.txt 2>&1; cat gpurun_out/r2_train_bench.txt
