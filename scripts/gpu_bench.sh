python -m pytest tests/test_cav_wil_gpu.py tests/test_full_size_gpu.py tests/test_robustness_gpu.py tests/test_files_gpu.py -x -q 2>&1 | tail -2
python scripts/exp/wil_time.py
