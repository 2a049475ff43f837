set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=5 > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 > gpurun_out/r02p_bench2.json 2> gpurun_out/r02p_bench2.err; echo "bench2 rc=$?" >> gpurun_out/r02p_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --impl reference > gpurun_out/r02p_bench2_ref.json 2>> gpurun_out/r02p_bench2.err
tail -8 gpurun_out/r02p_pytest.log; tail -c 500 gpurun_out/r02p_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r02p_bench2.json')); print(d['value'], d['e2e']['value'], d['clocks']); print(json.dumps(d['train_iter'], indent=1))"; cat gpurun_out/r02p_bench2_ref.json | cut -c1-600
