set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02n_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_wgrad.py -m gpu -q -x > gpurun_out/r02n_pytest_wgrad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest_wgrad.log
timeout 120 python tools/bench_wgrad.py > gpurun_out/wgrad_r02.json 2> gpurun_out/r02n.err
timeout 600 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_entrypoint.py -m gpu -q > gpurun_out/r02n_pytest_ppo.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest_ppo.log
LHW_WGRAD_KERNEL=0 python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02n_train_fp64_off.json 2>> gpurun_out/r02n.err
python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02n_train_fp64_on.json 2>> gpurun_out/r02n.err
python tools/bench_train_iter.py 4096 400 21845 32 > gpurun_out/r02n_train_fp32_on.json 2>> gpurun_out/r02n.err
python tools/prof_update.py 21845 > gpurun_out/r02n_update.log 2>&1
tail -5 gpurun_out/r02n_pytest_wgrad.log; tail -5 gpurun_out/r02n_pytest_ppo.log; cat gpurun_out/wgrad_r02.json; cat gpurun_out/r02n_train_fp64_off.json gpurun_out/r02n_train_fp64_on.json gpurun_out/r02n_train_fp32_on.json; tail -22 gpurun_out/r02n_update.log | cut -c1-150; tail -c 600 gpurun_out/r02n.err
