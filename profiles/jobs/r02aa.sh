# r02aa: compute-sanitizer over small renders of every kernel shape
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_small.py > gpurun_out/r02aa_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02aa_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_small.py > gpurun_out/r02aa_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02aa_racecheck.log | cut -c1-200
