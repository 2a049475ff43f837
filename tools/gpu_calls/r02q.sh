set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02q_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -f -o gpurun_out/prof_r2q_h164 python tools/prof_one.py 64 4096 62 0.223 h1 > gpurun_out/r02q_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -f -o gpurun_out/prof_r2q_terrain64 python tools/prof_one.py 64 4096 62 0.223 jvrc_walk_terrain >> gpurun_out/r02q_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_mw -s 60 -c 1 -f -o gpurun_out/prof_r2q_h132 python tools/prof_one.py 32 4096 62 0.223 h1 >> gpurun_out/r02q_ncu.log 2>&1
tail -5 gpurun_out/r02q_ncu.log; ls -la gpurun_out/prof_r2q_*
