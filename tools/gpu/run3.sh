mkdir -p gpurun_out
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prove.py tests/test_gpu_multirank.py -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
python bench.py --no-north-star --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_nttmul2.so python -m pytest tests/test_gpu_primitives.py -x -q -k ntt > gpurun_out/r2q_pytest_nttmul2.log 2>&1; tail -2 gpurun_out/r2q_pytest_nttmul2.log
TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_nttmul2.so python bench.py --no-north-star --no-cpu-baseline --no-sweep --steps 10 --warmup 3 > gpurun_out/r2q_bench_nttmul2.json 2> gpurun_out/r2q_bench_nttmul2.err
python - <<'P'
import json
for f in ["r2q_bench","r2q_bench_nttmul2"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"])
        print("   ", d["standalone"])
        for sw in d.get("sweeps") or []:
            if sw.get("sweep")=="msm" and sw.get("log_n") in (16,20,24): print("   ", {k:sw[k] for k in ("log_n","scalars","ms","mpts_per_s","phases_ms")})
    except Exception as e:
        print(f, "ERR", e)
P
