set -x
mkdir -p gpurun_out
# launch list of the final code state (cold-cache, serialised: only the SHARES are meaningful)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1y_launches_prove_2p20.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
# memcheck of one small end-to-end pass: build, prove, verify, wire formats
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import __graft_entry__ as g
g.smoke()
from typlonk_b200 import ffi
from typlonk_b200.kzg import Srs
from oracle.pyoracle import rng
ctx = ffi.Context(0)
s = Srs.from_secret(ctx, rng.fr_rand_stream(1, 1)[0], 509)
raw = s.to_bytes()
b = Srs.from_bytes(ctx, raw, 2)
assert b.handle.download() == s.handle.download()
print('wire ok')
" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"
tail -6 gpurun_out/memcheck.log
