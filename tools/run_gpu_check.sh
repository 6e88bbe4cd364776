set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_msm_affine.py -m gpu -x -q ) 2>&1 | tail -6 | cut -c1-200
for V in "" inv5 inv5b3; do
  if [ -n "$V" ]; then export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_$V.so; else unset TYPLONK_B200_LIB; fi
  for L in 20 22; do TP_MSM_AFFINE=1 timeout 300 python tools/msm_bench.py --log-n $L; done
done
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_inv5.so
TP_MSM_AFFINE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_affine -s 2 -c 1 -o gpurun_out/prof_afc_r1q -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_afc.log 2>&1
ls -la gpurun_out | tail -3
