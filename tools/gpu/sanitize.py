"""compute-sanitizer target: the code paths added in round 2 (radix-4 NTT rounds, MSM pipe, sharded digit pass) at small
sizes, each compared with the default path.  Run under `compute-sanitizer --tool memcheck|racecheck`."""
import numpy as np

from typlonk_b200 import field as F, synthetic
from typlonk_b200.ffi import Context

ctx = Context(0)
g = F.fr_to_bytes(7)
for log_n in (3, 4, 7, 10, 11, 13):
    n = 1 << log_n
    rs = np.random.RandomState(log_n)
    raw = rs.randint(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    data = raw.tobytes()
    outs = []
    for radix in (3, 2):
        ctx.set_option("ntt_radix_log", radix)
        outs.append((ctx.ntt(data, log_n), ctx.ntt(data, log_n, inverse=True), ctx.ntt(data, log_n, coset_mont=g)))
    assert outs[0] == outs[1], log_n
print("ntt ok")
log_n = 9
n = 1 << log_n
circuit = synthetic.mul_chain_direct(ctx, log_n)
cols = [F.fr_vec_to_bytes(c) for c in synthetic.mul_chain_witness(n - 3, n)]
want = circuit.handle.prove(cols, bytes(32 * n))
ctx.set_option("msm_pipeline", 2)
ctx.set_option("msm_pipe_min_log", 0)
assert circuit.handle.prove(cols, bytes(32 * n)) == want
ctx.set_option("msm_acc_staged", 1)
assert circuit.handle.prove(cols, bytes(32 * n)) == want
ctx.set_option("msm_pipeline", 0)
ctx.set_option("msm_acc_staged", 0)
print("pipe ok")
grp = Context.multi([0, 0, 0, 0])
gc = synthetic.mul_chain_direct(grp, log_n)
assert gc.handle.prove(cols, bytes(32 * n)) == want
grp.set_option("msm_pipeline", 2)
grp.set_option("msm_pipe_min_log", 0)
assert gc.handle.prove(cols, bytes(32 * n)) == want
grp.close()
print("sharded ok")
