mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
tail -c 3000 gpurun_out/bench_1080p.json; tail -5 gpurun_out/bench_1080p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --quick > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:blurKernel -s 6 -c 5 -o gpurun_out/prof_blur_r1 python bench.py --steps 1 --quick > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
