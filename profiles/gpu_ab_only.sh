bash profiles/gpu_ab.sh 2>&1 | grep -v "e2e ms"
