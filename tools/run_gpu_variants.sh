for lg in 22 24; do
for v in 0 4 6 8; do
  export TP_MSM_SCATTER_PART=$v
  timeout 300 python -m typlonk_b200.sweep --msm $lg --ntt "" --reps 3 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lg $lg part_lp[$v]', round(d['ms'],3), {k:round(x,3) for k,x in d['phases_ms'].items()})"
done
done
