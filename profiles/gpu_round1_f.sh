# launch list + full captures of blur (octave 0), descriptor, orientation, extrema, gradient
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --quick > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blurKernel -s 1 -c 5 -o gpurun_out/prof_blur python bench.py --steps 1 --quick > gpurun_out/ncu_full.log 2>&1
for k in descriptorKernel orientationKernel extremaMaskKernel gradientKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_$k python bench.py --steps 1 --quick > gpurun_out/ncu_$k.log 2>&1
done
ls gpurun_out | head -30
