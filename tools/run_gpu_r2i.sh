set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2i_bench_n$N.json").read().strip().splitlines()[-1])
print(round(d["value"], 3), round(d["e2e"]["value"], 3), d["parity"]["digest_ok"], d["parity"].get("ranks_agree"), d["gpu_launches"])
print("   ", json.dumps(d["phases_ms_per_step"]), json.dumps(d["standalone"])[:160])
if d.get("north_star"): print("   ns", json.dumps(d["north_star"])[:700])
PY
tail -3 gpurun_out/r2i_bench_n$N.err
