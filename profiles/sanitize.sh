# compute-sanitizer passes over the smoke test (320x240 frame through the whole path)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python __graft_entry__.py --smoke > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
