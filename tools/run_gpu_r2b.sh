# N-GPU check of the library-owned NCCL paths: single-process device group (C++ host) and torchrun bench
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
g++ -std=c++17 -O1 -Iinclude -o /tmp/mirror_test tests/cpp/mirror_test.cpp -Ltyplonk_b200/lib -ltyplonk_b200 -Wl,-rpath,$PWD/typlonk_b200/lib
timeout 600 /tmp/mirror_test multi $DEVS > gpurun_out/r2b_mirror_multi_n$N.log 2>&1; echo "mirror multi rc=$?"
grep -c "^PROOF" gpurun_out/r2b_mirror_multi_n$N.log; grep -v "^PROOF\|^ok" gpurun_out/r2b_mirror_multi_n$N.log | tail -5
python - <<PY
import json
seen = {l.split()[1]: l.split()[2] for l in open("gpurun_out/r2b_mirror_multi_n$N.log") if l.startswith("PROOF ")}
gold = json.load(open("tests/golden/proofs.json"))
print("golden match:", all(seen.get(k) == v["proof_hex"] for k, v in gold.items()), len(seen))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2b_bench_n$N.json; tail -5 gpurun_out/r2b_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_bench_n$N.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "parity", "phases_ms_per_step", "standalone", "north_star", "gpu_launches"):
    print(k, json.dumps(d.get(k))[:700])
PY
