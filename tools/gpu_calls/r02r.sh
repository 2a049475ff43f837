set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02r_build.log 2>&1
LHW_TENSORBOARD=0 timeout 400 python run_experiment.py train --env jvrc_walk --num-procs 4096 --n-itr 80 --seed 0 --eval-freq 1000 --logdir /tmp/lhw_r02r_walk > gpurun_out/r02r_train_walk.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_train_walk.log
LHW_TENSORBOARD=0 timeout 300 python run_experiment.py train --env h1 --num-procs 4096 --n-itr 40 --seed 0 --eval-freq 1000 --logdir /tmp/lhw_r02r_h1 > gpurun_out/r02r_train_h1.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_train_h1.log
grep -E "Mean Eprew|Mean Eplen|Total time" gpurun_out/r02r_train_walk.log | awk 'NR%30<3' | head -40; tail -3 gpurun_out/r02r_train_walk.log; grep -E "Mean Eprew|Mean Eplen" gpurun_out/r02r_train_h1.log | tail -4; tail -2 gpurun_out/r02r_train_h1.log
