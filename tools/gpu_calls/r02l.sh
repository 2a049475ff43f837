set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02l_build.log 2>&1
python -m pytest tests -m gpu -q --durations=5 -p no:randomly > gpurun_out/r02l_pytest_natural.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest_natural.log
for k in gae_kernel adv_apply_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_r2l_$k python tools/bench_ppo_kernels.py >> gpurun_out/r02l_ncu.log 2>&1
done
python bench.py --workload h1 --no-cpu-baseline --no-train-iter --no-extras --steps 200 > gpurun_out/r02l_bench_h1.json 2> gpurun_out/r02l_bench.err
tail -8 gpurun_out/r02l_pytest_natural.log; tail -c 400 gpurun_out/r02l_bench_h1.json
