set -x
python -m pytest tests -m gpu -q --durations=8 -k "not n_rank and not uneven" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
python tools/bench_ppo_kernels.py > gpurun_out/ppo_kernels_r02.json 2> gpurun_out/r02d_ppo.err
python tools/bench_train_iter.py 4096 400 21845 32 > gpurun_out/r02d_train_fp32.json 2>> gpurun_out/r02d_ppo.err
python tools/bench_train_iter.py 4096 400 21845 32 tf32 > gpurun_out/r02d_train_fp32_tf32.json 2>> gpurun_out/r02d_ppo.err
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -o gpurun_out/prof_r2d_walk64 python tools/prof_one.py 64 4096 62 0.223 jvrc_walk > gpurun_out/r02d_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -o gpurun_out/prof_r2d_walk32 python tools/prof_one.py 32 4096 62 0.223 jvrc_walk >> gpurun_out/r02d_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -o gpurun_out/prof_r2d_step64 python tools/prof_one.py 64 4096 62 0.223 jvrc_step >> gpurun_out/r02d_ncu.log 2>&1
ncu --set full --clock-control none -k 'regex:gae_kernel|adv_apply|adv_from_gae|gather_kernel|exchange_reduce|clip_adam_pair|exchange_finish|sumsq_kernel|clip_adam_dev' -c 14 -s 40 -o gpurun_out/prof_r2d_ppo python tools/bench_ppo_kernels.py >> gpurun_out/r02d_ncu.log 2>&1
python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?" >> gpurun_out/r02d_bench.err
tail -12 gpurun_out/r02d_pytest.log; cat gpurun_out/ppo_kernels_r02.json; cat gpurun_out/r02d_train_fp32.json gpurun_out/r02d_train_fp32_tf32.json; tail -c 400 gpurun_out/r02d_bench.err; head -c 300 gpurun_out/r02d_bench.json
