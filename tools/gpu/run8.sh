mkdir -p gpurun_out
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_ntt2.so
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2w_ntt_b2_2blk.jsonl 2>/dev/null
cat gpurun_out/r2w_ntt_b2_2blk.jsonl | cut -c1-160
