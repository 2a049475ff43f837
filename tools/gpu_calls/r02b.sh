set -x
python -m pytest tests -m gpu -q --durations=8 -s -k "not n_rank and not uneven" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
python bench.py --impl reference > gpurun_out/r02b_bench_ref.json 2>> gpurun_out/r02b_bench.err
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -o gpurun_out/prof_r2b_walk64 python tools/prof_one.py 64 4096 62 0.223 jvrc_walk > gpurun_out/r02b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -o gpurun_out/prof_r2b_walk32 python tools/prof_one.py 32 4096 62 0.223 jvrc_walk >> gpurun_out/r02b_ncu.log 2>&1
tail -5 gpurun_out/r02b_pytest.log; tail -c 600 gpurun_out/r02b_bench.err; head -c 1500 gpurun_out/r02b_bench.json
