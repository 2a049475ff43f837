set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_ppo.py tests/test_gpu_h1.py::test_ppo_trains_on_the_h1_environment tests/test_gpu_entrypoint.py tests/test_gpu_parity.py -m gpu -q --durations=8 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02c_bench2.json 2> gpurun_out/r02c_bench2.err; echo "bench2 rc=$?" >> gpurun_out/r02c_bench2.err
LHW_WARPS_PER_BLOCK=16 python tools/quick_bench.py jvrc_walk 4096,32768 > gpurun_out/r02c_wpb16.log 2>&1
tail -15 gpurun_out/r02c_pytest.log; tail -c 800 gpurun_out/r02c_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench2.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['train_iter'], indent=1))"; cat gpurun_out/r02c_wpb16.log
