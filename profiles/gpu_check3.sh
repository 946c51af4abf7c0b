SIFTCUDA_BANDS=3 timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 800 python -m pytest tests -m gpu -x -q -k "1080p or butterfly" 2>&1 | tail -2
bash profiles/gpu_ab.sh 2>&1
