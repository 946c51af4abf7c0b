# compute-sanitizer passes over the smoke test (320x240 frame through the whole path) and a 1080p frame
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python __graft_entry__.py --smoke > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
for tool in memcheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python profiles/sanitize_1080p.py > gpurun_out/sanitize_1080p_$tool.log 2>&1
  echo "== 1080p $tool: $(grep -E 'ERROR SUMMARY|1080p ok' gpurun_out/sanitize_1080p_$tool.log | tr '\n' ' ')"
done
