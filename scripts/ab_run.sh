mkdir -p gpurun_out
python -m pytest tests/test_dataops_gpu.py tests/test_transforms_gpu.py -q 2>&1 | tail -5
