timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SIFTCUDA_BANDS=4 timeout 800 python -m pytest tests -m gpu -x -q -k "1080p" 2>&1 | tail -2
bash profiles/gpu_ab.sh 2>&1
