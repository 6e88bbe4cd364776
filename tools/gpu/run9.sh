mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu/sanitize.py > gpurun_out/r2x_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2x_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/gpu/sanitize.py > gpurun_out/r2x_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2x_racecheck.log
