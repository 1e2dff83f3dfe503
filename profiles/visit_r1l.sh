timeout 500 python profiles/dbg_pair.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15
