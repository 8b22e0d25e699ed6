mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_shapes or tc" > gpurun_out/edge_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/edge_pytest.txt
