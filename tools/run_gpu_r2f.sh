set -x
N=${1:-8}
mkdir -p gpurun_out
run() { # tag extra-env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/r2f_${tag}_n$N.json 2> gpurun_out/r2f_${tag}_n$N.err; echo "$tag rc=$?"
}
EXTRA="" run main X=1
EXTRA="--no-north-star" run noside TP_NO_SIDE_STREAM=1
python - <<PY
import json
for tag in ("main", "noside"):
    d = json.loads(open("gpurun_out/r2f_%s_n$N.json" % tag).read().strip().splitlines()[-1])
    print(tag, round(d["value"], 3), round(d["e2e"]["value"], 3), d["parity"], d["gpu_launches"])
    print("   ", json.dumps(d["phases_ms_per_step"]), json.dumps(d["standalone"])[:200])
    if d.get("north_star"): print("   ns", json.dumps(d["north_star"])[:700])
PY
tail -3 gpurun_out/r2f_main_n$N.err
