set -x
mkdir -p gpurun_out
for V in "" nttcall; do
  if [ -n "$V" ]; then export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_$V.so; else unset TYPLONK_B200_LIB; fi
  timeout 600 python -m typlonk_b200.sweep --msm 16,18 --ntt 12,16,18,20,22,24 2>&1 | cut -c1-220
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_ms_per_step'])"
done
