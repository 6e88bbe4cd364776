mkdir -p gpurun_out
python -m pytest tests/test_gpu_primitives.py -x -q -k "ntt" > gpurun_out/r2r_pytest.log 2>&1; tail -2 gpurun_out/r2r_pytest.log
python bench.py --no-north-star --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_nttinl.so
python -m pytest tests/test_gpu_primitives.py -x -q -k "ntt" > gpurun_out/r2r_pytest_inl.log 2>&1; tail -2 gpurun_out/r2r_pytest_inl.log
python bench.py --no-north-star --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2r_bench_inl.json 2> gpurun_out/r2r_bench_inl.err
python - <<'P'
import json
for f in ["r2r_bench","r2r_bench_inl"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"])
        print("   ", d["standalone"])
        for sw in d.get("sweeps") or []:
            if sw.get("sweep")!="msm": print("   ", sw)
    except Exception as e:
        print(f, "ERR", e)
P
