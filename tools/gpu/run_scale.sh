mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err
tail -2 gpurun_out/r2y_bench_n$N.err | cut -c1-300
python - <<P
import json
d=json.loads(open("gpurun_out/r2y_bench_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"], d["phases_ms_per_step"])
ns=d["north_star"]; print({k:ns[k] for k in ns if k not in ("parity",)}, ns["parity"]["digest_ok"])
print(d["standalone"]); print(d.get("ab"))
P
