"""GPU, several ranks on ONE device (gloo carries the collectives through host memory): the
multi-GPU plan of SURVEY.md 8(e) exactly as bench.py runs it under torchrun -- every MSM sharded
by contiguous point range with the 144-byte partial points all-gathered, the quotient sharded by
coset of the 4n domain with one device broadcast per coset -- must produce the proof bytes of the
single-rank prover (and of the C++ oracle).  world = 2, 3 and 4 cover 2 / uneven / 1 coset per rank."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, log_n, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from typlonk_b200 import field as F, synthetic
        from typlonk_b200.ffi import Context, DeviceView
        dev = torch.device("cuda", 0)
        st = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(st)
        ctx = Context(0, st.cuda_stream)

        def allgather(data: bytes) -> bytes:
            send = torch.frombuffer(bytearray(data), dtype=torch.uint8)
            recv = torch.empty(world * len(data), dtype=torch.uint8)
            dist.all_gather_into_tensor(recv, send)
            return recv.numpy().tobytes()

        def bcast(ptr: int, nbytes: int, root: int):
            t = torch.as_tensor(DeviceView(ptr, nbytes), device=dev)
            host = t.cpu()                      # synchronises with the ctx stream (current stream)
            dist.broadcast(host, src=root)
            if rank != root:
                t.copy_(host)
                torch.cuda.current_stream().synchronize()

        ctx.set_shard(rank, world, allgather)
        ctx.set_broadcast(bcast)
        n = 1 << log_n
        circuit = synthetic.mul_chain_direct(ctx, log_n)
        cols = synthetic.mul_chain_witness(n - 3, n)
        col_bytes = [F.fr_vec_to_bytes(c) for c in cols]
        proof = circuit.handle.prove(col_bytes, bytes(32 * n))
        # the short public-input vector of prove()'s caller: three sliced columns + one whole 32-byte vector
        short = circuit.handle.prove_inputs(col_bytes, bytes(32))
        if short != proof:
            proof = "error: tp_prove_inputs and tp_prove disagree on rank %d" % rank
        q.put((rank, proof))
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: %r\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world,log_n", [(2, 10), (3, 8), (4, 12)])
def test_sharded_prover_matches_single_rank_and_oracle(ctx, world, log_n):
    from oracle import coracle
    from typlonk_b200 import field as F, synthetic
    n = 1 << log_n
    circuit = synthetic.mul_chain_direct(ctx, log_n)
    cols = synthetic.mul_chain_witness(n - 3, n)
    single = circuit.handle.prove([F.fr_vec_to_bytes(c) for c in cols], bytes(32 * n))
    tau_b, sel, perm, ocols, pi = coracle.mul_chain_inputs(log_n)
    oc = coracle.Circuit(tau_b, sel, perm, n)
    assert oc.prove(ocols, pi) == single
    oc.close()

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29700 + (os.getpid() % 1000) + world
    procs = [mpc.Process(target=_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert results[r] == single, "rank %d: %s" % (r, results[r] if isinstance(results[r], str) else "proof differs")
