mkdir -p gpurun_out
python scripts/sanitize_small.py > gpurun_out/r2af_plain.log 2>&1; echo "plain rc=$?"; tail -6 gpurun_out/r2af_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py 3,6,17 > gpurun_out/r2af_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -h "ERROR SUMMARY\|Invalid\|out of bounds\|L = \|ft rows" gpurun_out/r2af_memcheck.log | head -20
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_small.py 6 > gpurun_out/r2af_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -h "RACECHECK SUMMARY\|hazard\|L = \|ft rows" gpurun_out/r2af_racecheck.log | head -20
