mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py tests/test_gpu_primitives.py -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -5 gpurun_out/r2o_pytest.log
python bench.py --no-sweep --no-north-star --no-cpu-baseline --steps 10 --warmup 3 --ab msm_acc_staged=1 --ab msm_pipeline=2,msm_acc_staged=1 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"])
for a in d["ab"]: print(a)
P
