set -x
python -m pytest tests/test_gpu_ppo.py tests/test_gpu_entrypoint.py -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
LHW_FUSED_LOSS=0 python tools/prof_update.py 21845 > gpurun_out/r02f_update_unfused.log 2>&1
LHW_FUSED_LOSS=1 python tools/prof_update.py 21845 > gpurun_out/r02f_update_fused.log 2>&1
python tools/bench_train_iter.py 4096 400 21845 32 > gpurun_out/r02f_train_fp32.json 2> gpurun_out/r02f.err
python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02f_train_fp64.json 2>> gpurun_out/r02f.err
for k in gae_kernel gather_kernel exchange_reduce_kernel clip_adam_pair_kernel exchange_finish_kernel sumsq_kernel clip_adam_dev_kernel adv_apply_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 4 -c 1 -o gpurun_out/prof_r2f_$k python tools/bench_ppo_kernels.py >> gpurun_out/r02f_ncu.log 2>&1
done
ncu --set full --clock-control none -k regex:ppo_loss_kernel -s 3 -c 1 -o gpurun_out/prof_r2f_ppo_loss_kernel python tools/prof_update.py 21845 >> gpurun_out/r02f_ncu.log 2>&1
tail -6 gpurun_out/r02f_pytest.log; tail -18 gpurun_out/r02f_update_unfused.log; tail -18 gpurun_out/r02f_update_fused.log; cat gpurun_out/r02f_train_fp32.json gpurun_out/r02f_train_fp64.json
