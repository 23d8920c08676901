cd /root/repo
mkdir -p gpurun_out
which compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done
