set -x
mkdir -p gpurun_out
for ch in 0 96 128 192; do
  TP_MSM_CHUNK=$ch timeout 300 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --no-north-star > gpurun_out/r2e_chunk$ch.json 2>> gpurun_out/r2e.err
done
python - <<PY
import json
for ch in (0, 96, 128, 192):
    d = json.loads(open("gpurun_out/r2e_chunk%d.json" % ch).read().strip().splitlines()[-1])
    print(ch, round(d["value"], 3), d["parity"]["digest_ok"], json.dumps(d["phases_ms_per_step"]))
PY
tail -3 gpurun_out/r2e.err
