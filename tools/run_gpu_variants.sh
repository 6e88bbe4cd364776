mkdir -p gpurun_out
run() {
  timeout 200 python bench.py --log-n 17 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],3), d['phases_ms_per_step'], d['roofline_int']['window_bits'], d['roofline_int']['windows'])"
}
run default
TP_MSM_WSUM_MIN=8192 TP_MSM_WSUM_S1=8 run "min8192,s1=8"
TP_MSM_WSUM_MIN=4096 TP_MSM_WSUM_S1=8 run "min4096,s1=8"
TP_MSM_WSUM_MIN=8192 TP_MSM_WSUM_S1=16 run "min8192,s1=16"
TP_MSM_WSUM_MIN=32768 TP_MSM_WSUM_S1=4 run "min32768,s1=4"
TP_MSM_WSUM_MIN=2048 TP_MSM_WSUM_S1=4 run "min2048,s1=4"
TP_MSM_WSUM_MIN=65536 run "no levels"
TP_MSM_C=15 run "c=15"
TP_MSM_C=17 run "c=17"
TP_MSM_C=18 run "c=18"
