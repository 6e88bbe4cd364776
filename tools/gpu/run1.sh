set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
python bench.py --no-sweep --no-north-star --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2n_bench_pipe.json 2> gpurun_out/r2n_bench_pipe.err
TP_MSM_PIPELINE=0 python bench.py --no-sweep --no-north-star --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2n_bench_nopipe.json 2> gpurun_out/r2n_bench_nopipe.err
python - <<'P'
import json
for f in ["pipe","nopipe"]:
    try:
        d=json.loads(open("gpurun_out/r2n_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
P
