set -x
python -m pytest tests/test_gpu_ppo.py -m gpu -q -k "two_stream or rollout" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
python tools/ab_variants.py run > gpurun_out/r02g_ab.log 2>&1
python tools/bench_train_iter.py 4096 400 21845 32 > gpurun_out/r02g_train_fp32.json 2> gpurun_out/r02g.err
python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02g_train_fp64.json 2>> gpurun_out/r02g.err
LHW_ROLLOUT_SPLIT=0 python tools/bench_train_iter.py 4096 400 21845 64 > gpurun_out/r02g_train_fp64_nosplit.json 2>> gpurun_out/r02g.err
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?" >> gpurun_out/r02g_bench.err
tail -4 gpurun_out/r02g_pytest.log; tail -12 gpurun_out/r02g_ab.log; cat gpurun_out/r02g_train_fp32.json gpurun_out/r02g_train_fp64.json gpurun_out/r02g_train_fp64_nosplit.json; tail -c 300 gpurun_out/r02g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['issue'], d['extras'].get('rollout_with_policy_env_steps_per_s_per_gpu')); print(d['train_iter']['fp64'], d['train_iter']['fp32'])"
