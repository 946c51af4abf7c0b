for parts in 0 4; do
  SIFTCUDA_DESC_PARTS=$parts ncu --set full --clock-control none --import-source on -k regex:descriptorKernel -s 2 -c 1 -o gpurun_out/prof_desc_parts$parts python bench.py --steps 2 --quick > gpurun_out/ncu_desc_parts$parts.log 2>&1
done
ls -la gpurun_out/prof_desc_parts*
