# ncu evidence: launch list of one proof + full sets of the top kernels (1 GPU); reports are exported to CSV on the
# box (gpurun brings back at most 64 MiB) and only the accumulate report itself is kept
set -x
mkdir -p gpurun_out /tmp/rep
B="python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-north-star"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_prove_2p20.csv $B > /tmp/rep/ncu_bench.log 2>&1
full() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/rep/r2d_$1 $B > /tmp/rep/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/rep/r2d_$1.ncu-rep --page raw --csv > gpurun_out/r2d_$1_raw.csv 2>/dev/null
  ncu -i /tmp/rep/r2d_$1.ncu-rep --page details --csv > gpurun_out/r2d_$1_details.csv 2>/dev/null
}
full accumulate '^k_msm_accumulate$' 3 3
ncu -i /tmp/rep/r2d_accumulate.ncu-rep --page source --csv --print-source sass > /tmp/rep/acc_source.csv 2>/dev/null; gzip -c /tmp/rep/acc_source.csv > gpurun_out/r2d_accumulate_source_sass.csv.gz
full wsum 'k_msm_wsum_level' 9 9
full ntt 'k_ntt_r8' 60 8
ncu -i /tmp/rep/r2d_ntt.ncu-rep --page source --csv --print-source sass > /tmp/rep/ntt_source.csv 2>/dev/null; gzip -c /tmp/rep/ntt_source.csv > gpurun_out/r2d_ntt_source_sass.csv.gz
full sort 'k_msm_digits|k_msm_scatter' 6 6
full tail 'k_msm_masked_sum|k_msm_pair_fixup' 9 9
du -sh gpurun_out
