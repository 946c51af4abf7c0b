# full GPU test suite + A/B bench lines from profiles/ab_variants.txt
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/gpu_ab.sh
