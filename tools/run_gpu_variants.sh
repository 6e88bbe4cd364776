mkdir -p gpurun_out
for v in "" _prefetch _prefetch2 ""; do
  export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200$v.so
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant[$v]', round(d['value'],3), round(d['e2e']['value'],3), d['phases_ms_per_step'], d['proof_sha256'])"
done
