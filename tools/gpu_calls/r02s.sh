set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02s_build.log 2>&1
LHW_TENSORBOARD=0 timeout 400 python run_experiment.py train --env jvrc_walk --num-procs 4096 --n-itr 40 --seed 0 --eval-freq 40 --logdir /tmp/lhw_r02s_walk > gpurun_out/r02s_train_walk.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_train_walk.log
mkdir -p gpurun_out/trained_walk; find /tmp/lhw_r02s_walk -name "*.pt" -exec cp {} gpurun_out/trained_walk/ \; ; find /tmp/lhw_r02s_walk -name "*.pkl" -exec cp {} gpurun_out/trained_walk/ \; ; ls -la gpurun_out/trained_walk
grep -E "Mean Eplen" gpurun_out/r02s_train_walk.log | tail -2; grep -A1 EVALUATE gpurun_out/r02s_train_walk.log | tail -2
