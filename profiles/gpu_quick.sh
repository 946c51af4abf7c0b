timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "butterfly or synthetic or 1080p" 2>&1 | tail -2
bash profiles/gpu_ab.sh 2>&1
