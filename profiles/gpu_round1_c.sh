# full ncu captures of the non-blur hot kernels (first launch of each)
mkdir -p gpurun_out
for k in descriptorKernel orientationKernel extremaMaskKernel gradientKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_$k python bench.py --steps 1 --quick > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
