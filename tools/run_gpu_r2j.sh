set -x
N=${1:-4}
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 6 --warmup 3 --log-n 22 --no-north-star > gpurun_out/r2j_bench22_n${N}_$i.json 2> gpurun_out/r2j_$i.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2j_bench22_n${N}_$i.json").read().strip().splitlines()[-1])
print(round(d["value"], 3), round(d["e2e"]["value"], 3), d["parity"]["digest_ok"], json.dumps(d["phases_ms_per_step"]))
PY
done
