set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02t_build.log 2>&1
python -m pytest tests/test_gpu_parity_shipped.py -m gpu -q -s -k "trained" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.log
tail -12 gpurun_out/r02t_pytest.log
