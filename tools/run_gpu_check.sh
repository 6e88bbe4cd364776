set -x
mkdir -p gpurun_out
tools/ubench/fqmul_forms > gpurun_out/fqmul_forms.txt 2>&1; cat gpurun_out/fqmul_forms.txt
for v in "" _base _kar _sqronly _lazyonly; do
  export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200$v.so
  echo "variant=$v" >> gpurun_out/msm_variants.txt
  timeout 300 python -c "
from typlonk_b200.ffi import Context
c=Context(0); print('selftest failures', c.selftest())" >> gpurun_out/msm_variants.txt 2>&1
  timeout 300 python tools/msm_bench.py --log-n 20 --reps 5 >> gpurun_out/msm_variants.txt 2>&1
done
cat gpurun_out/msm_variants.txt
