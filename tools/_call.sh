timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_online.py -x -q > gpurun_out/r2_pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_train.log; tail -15 gpurun_out/r2_pytest_train.log
timeout 300 python tools/bench_train.py > gpurun_out/r2_train_bench.txt 2>&1; cat gpurun_out/r2_train_bench.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "mlse or channel or 4100 or near_ties or config1" > gpurun_out/r2_pytest_new.log 2>&1; tail -15 gpurun_out/r2_pytest_new.log
