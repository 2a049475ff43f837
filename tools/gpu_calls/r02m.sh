set -x
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02m_build.log 2>&1
python -m pytest tests/test_gpu_ppo_edges.py -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -40 gpurun_out/r02m_pytest.log
