for lg in 22 24; do
for v in 1 2 4 8 16; do
  export TP_MSM_SCATTER_SLICES=$v
  timeout 300 python -m typlonk_b200.sweep --msm $lg --ntt "" --reps 3 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lg $lg slices[$v]', round(d['ms'],3), {k:round(x,3) for k,x in d['phases_ms'].items()})"
done
done
