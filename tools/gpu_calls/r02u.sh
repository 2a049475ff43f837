set -x
python bench.py --no-train-iter --no-cpu-baseline --steps 100 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?" >> gpurun_out/r02u_bench.err
tail -c 400 gpurun_out/r02u_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02u_bench.json')); print(d['value'], d['extras'])"
