mkdir -p gpurun_out
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2z_ntt_default.jsonl 2>/dev/null
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_lazy.so
python -m pytest tests/test_gpu_primitives.py -x -q -k "ntt" 2>&1 | tail -2
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2z_ntt_lazy.jsonl 2>/dev/null
python - <<'P'
import json
a=[json.loads(l) for l in open("gpurun_out/r2z_ntt_default.jsonl")]
b=[json.loads(l) for l in open("gpurun_out/r2z_ntt_lazy.jsonl")]
for x,y in zip(a,b): print(x["sweep"], x["log_n"], x["ms"], y["ms"], x["wide_mul_frac"], y["wide_mul_frac"])
P
