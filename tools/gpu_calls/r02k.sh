set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02k_smoke.log
python -m pytest tests -m gpu -q --durations=8 -k "not n_rank and not uneven" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?" >> gpurun_out/r02k_bench.err
python bench.py --impl reference > gpurun_out/r02k_bench_ref.json 2>> gpurun_out/r02k_bench.err
for w in jvrc_step h1 jvrc_walk_terrain; do python bench.py --workload $w --no-cpu-baseline --no-train-iter --steps 200 > gpurun_out/r02k_bench_$w.json 2>> gpurun_out/r02k_bench.err; done
python bench.py --envs 1024 --workload jvrc_step --no-cpu-baseline --no-train-iter --no-extras --steps 200 > gpurun_out/r02k_bench_jvrc_step_1024.json 2>> gpurun_out/r02k_bench.err
for k in 3 4; do LHW_ROLLOUT_PARTS=$k python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02k_train_fp64_parts$k.json 2>> gpurun_out/r02k_bench.err; done
python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02k_train_fp64_parts2.json 2>> gpurun_out/r02k_bench.err
python tools/bench_ppo_kernels.py > gpurun_out/ppo_kernels_r02.json 2>> gpurun_out/r02k_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02k.csv python bench.py --steps 20 --warmup 3 --no-extras --no-train-iter --no-cpu-baseline > gpurun_out/r02k_ll.log 2>&1
tail -3 gpurun_out/r02k_smoke.log; tail -6 gpurun_out/r02k_pytest.log; tail -c 300 gpurun_out/r02k_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_bench*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['e2e']['value']), d.get('extras',{}).get('rollout_with_policy_env_steps_per_s_per_gpu'), (d.get('train_iter') or {}).get('fp64',{}).get('fps'))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/r02k_train*.json')):
    d=json.load(open(f)); print(f, round(d['fps_sampling_plus_optimisation']), d['sample_s'], d['optimize_s'])
d=json.load(open('gpurun_out/ppo_kernels_r02.json')); print({k:round(v['ms'],4) for k,v in d.items() if isinstance(v,dict)})
PY
