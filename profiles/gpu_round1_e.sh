mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:descriptorKernel -c 1 -o gpurun_out/prof_descriptorKernel python bench.py --steps 1 --quick > gpurun_out/ncu_desc.log 2>&1
tail -1 gpurun_out/ncu_desc.log
