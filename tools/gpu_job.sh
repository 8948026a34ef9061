mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "not size and not c2 and not c3" > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck.log | tail -4
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "assembly_rectangle or assembly_unstructured or morton or multigrid_preconditioner or lu_path or tau_not_one" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_racecheck.log | tail -6
