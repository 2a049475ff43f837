set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02o_smoke.log
ncu --set full --clock-control none --import-source on -k regex:wgrad_partial_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2o_wgrad_partial python tools/bench_wgrad.py > gpurun_out/r02o_ncu.log 2>&1
ncu --set full --clock-control none -k regex:wgrad_reduce_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2o_wgrad_reduce python tools/bench_wgrad.py >> gpurun_out/r02o_ncu.log 2>&1
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?" >> gpurun_out/r02o_bench.err
python tools/bench_train_iter.py 4096 400 21845 64 tf32 > gpurun_out/r02o_train_fp64_tf32.json 2>> gpurun_out/r02o_bench.err
tail -3 gpurun_out/r02o_smoke.log; tail -8 gpurun_out/r02o_pytest.log; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench.json')); print(d['value'], d['e2e']['value'], json.dumps(d['train_iter'])[:1500])"; cat gpurun_out/r02o_train_fp64_tf32.json; tail -c 300 gpurun_out/r02o_bench.err
