mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py 3,6,9,17 > gpurun_out/r2ag_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -h "ERROR SUMMARY\|Invalid\|out of bounds\|L = \|ft rows\|Error" gpurun_out/r2ag_memcheck.log | head -20
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_small.py 6,17 > gpurun_out/r2ag_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -h "RACECHECK SUMMARY\|hazard\|L = \|ft rows" gpurun_out/r2ag_racecheck.log | head -20
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python scripts/sanitize_small.py 6 > gpurun_out/r2ag_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -h "ERROR SUMMARY\|Uninitialized\|L = \|ft rows" gpurun_out/r2ag_initcheck.log | head -10
