timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash profiles/gpu_ab.sh 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --quick > /dev/null 2>&1
grep -E "gradientKernel|extremaMaskKernel|orientationKernel|descriptorKernel" gpurun_out/launches_tmp.csv | awk -F'","' '{print $5, $NF}' | head -12
