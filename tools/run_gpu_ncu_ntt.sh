mkdir -p gpurun_out
timeout 900 ncu --clock-control none -k regex:k_ntt_r8 -c 70 --section SpeedOfLight --section WarpStateStats --section SchedulerStats \
  --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section InstructionStats \
  -o gpurun_out/prof_ntt_r1y -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ntt.log 2>&1
tail -2 gpurun_out/ncu_ntt.log | cut -c1-200
ls -la gpurun_out/
