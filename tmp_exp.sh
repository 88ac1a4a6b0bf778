python -m pytest tests/test_parity_gpu.py tests/test_dist_gpu.py tests/test_extensions_gpu.py tests/test_scene_gpu.py -x -q -m gpu 2>&1 | tail -3
