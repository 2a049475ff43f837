set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8" --durations=5 > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 > gpurun_out/r02j_bench8.json 2> gpurun_out/r02j_bench8.err; echo "bench4 rc=$?" >> gpurun_out/r02j_bench8.err
tail -8 gpurun_out/r02j_pytest.log; tail -c 500 gpurun_out/r02j_bench8.err; python -c "
import json; d=json.load(open('gpurun_out/r02j_bench8.json')); print(d['value'], d['e2e']['value'], d['clocks']); print(json.dumps(d['train_iter'], indent=1)); print(d['extras'])"
