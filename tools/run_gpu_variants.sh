mkdir -p gpurun_out
for v in 1 2 4 8 0; do
  export TP_MSM_SCATTER_SLICES=$v
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('slices[$v]', round(d['value'],3), round(d['e2e']['value'],3), d['phases_ms_per_step'], d['proof_sha256'])"
done
