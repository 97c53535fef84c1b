mkdir -p gpurun_out
python -m pytest tests/test_sampling_gpu.py tests/test_transforms_gpu.py tests/test_dataops_gpu.py -q -s 2>&1 | grep -E "passed|failed|^FAILED|round trip|max \|x" | tail -70
