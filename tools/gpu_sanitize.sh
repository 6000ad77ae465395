#!/bin/bash
# compute-sanitizer passes over every kernel family (small sizes; the tools slow kernels down ~100x)
mkdir -p gpurun_out
g++ -std=c++17 -O1 -o /tmp/sanitize_main tests/cpp/sanitize_main.cpp -Lhades252_b200/lib -lhades_b200 -Wl,-rpath,$PWD/hades252_b200/lib || exit 1
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool /tmp/sanitize_main > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver OK|Error|hazard" gpurun_out/sanitize_$tool.log | head -5
done
