for tool in memcheck racecheck synccheck; do
  echo "== $tool" | tee -a gpurun_out/sanitizer.log
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2>&1 | tail -6 | tee -a gpurun_out/sanitizer.log
done
