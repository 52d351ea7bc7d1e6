"""CPU, world_size 2 and 3 over gloo: the N-column sharding plumbing (partition rule, single
broadcast of X, per-rank column slices, no reduction).  The per-rank compute is the checker
(oracle BaseTCSC on the slice's own TCSC) standing in for the CUDA kernel, so what is tested
is exactly the host logic bench.py runs under torchrun on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, K, N, s, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        ge.load_package()
        from ternary_spgemm_b200 import shard
        from oracle.pyoracle import Oracle
        orc = Oracle()
        W = orc.generate_sparse_matrix(K, N, s, 42)          # every rank can build its own slice
        b = np.linspace(-1, 1, N, dtype=np.float32)
        # only rank 0 holds the real X; the others start with garbage
        X = torch.from_numpy(orc.init_x(M, K, 7)) if rank == 0 else torch.full((M, K), -1e9)

        def compute(Xt, lo, hi):
            t = orc.tcsc(W[:, lo:hi])                         # == TCSC(W[:, lo:hi]) (shard semantics)
            return torch.from_numpy(orc.base_tcsc(Xt.numpy(), t, b[lo:hi]))

        Y_local, (lo, hi) = shard.sharded_spmm(X, N, compute)
        assert (lo, hi) == shard.shard_columns(N, world, rank) and Y_local.shape == (M, hi - lo)
        Y = shard.gather_columns(Y_local, N)
        Y_ref = orc.base_tcsc(orc.init_x(M, K, 7), orc.tcsc(W), b)
        ok = Y.shape == (M, N) and np.array_equal(Y.numpy(), Y_ref)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == float(world)
        open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 96), (3, 100), (2, 7)])
def test_sharded_path_over_gloo(tmp_path, world, N):
    M, K, s = 3, 64, 2
    mp.spawn(_worker, args=(world, _free_port(), M, K, N, s, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"rank{r}.ok" for r in range(world)]


def _worker_shm(rank, world, port, M, K, steps, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        ge.load_package()
        from ternary_spgemm_b200 import shard
        hx = shard.HostSharedX(M, K)
        ok = True
        for s in range(1, steps + 1):
            want = np.full((M, K), float(s), np.float32) + np.arange(K, dtype=np.float32)
            x = hx.next(want if rank == 0 else None)          # only rank 0 holds the step's X
            if rank == 1 and s % 3 == 0:
                import time
                time.sleep(0.01)                              # a slow reader must not be overrun
            ok = ok and x.shape == (M, K) and np.array_equal(x, want)
            hx.done()
        dist.barrier()
        hx.close()
        # a block that cannot fit in /dev/shm: EVERY rank gets the same RuntimeError (no rank is left
        # waiting in a collective), so callers can fall back together
        st = os.statvfs("/dev/shm")
        rows = (st.f_bavail * st.f_frsize) // (4 << 20) + 64          # x 2^20 floats x 2 buffers > free space
        try:
            shard.HostSharedX(int(rows), 1 << 20)
            ok = False
        except RuntimeError as e:
            ok = ok and "shared memory" in str(e)
        dist.barrier()
        # a reader that stops acknowledging (died, or forgot done()): the publisher gives up after
        # the timeout with an error instead of spinning forever; a reader whose publisher stopped too
        os.environ["TSG_SHM_TIMEOUT_S"] = "0.5"
        hz = shard.HostSharedX(M, K)
        try:
            if rank == 0:
                for s in range(1, 5):                         # the third publication needs an ack that never comes
                    hz.next(np.zeros((M, K), np.float32))
                    hz.done()
                ok = False
            else:
                hz.next(None)                                 # consume step 1, never call done(), then wait for a
                hz.next(None)                                 # step 2 that arrives and a step 3 that never does
                hz.next(None)
                ok = False
        except TimeoutError as e:
            ok = ok and "HostSharedX" in str(e)
        dist.barrier()
        hz.close()
        open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_host_shared_x_over_gloo(tmp_path, world):
    """X through host shared memory (the replication bench.py's e2e uses at N > 1): every rank
    sees every step's X, in order, and the publisher never overwrites a buffer still being read."""
    mp.spawn(_worker_shm, args=(world, _free_port(), 2, 40, 12, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"rank{r}.ok" for r in range(world)]


def _worker_hostcall(rank, world, port, M, K, N, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        ge.load_package()
        from ternary_spgemm_b200 import shard
        from oracle.pyoracle import Oracle
        orc = Oracle()
        W = orc.generate_sparse_matrix(K, N, 2, 5)
        b = np.linspace(-1, 1, N, dtype=np.float32)
        lo, hi = shard.shard_columns(N, world, rank)
        t = orc.tcsc(np.ascontiguousarray(W[:, lo:hi]))
        hc = shard.HostShardedCall(M, K, hi - lo, torch.device("cpu"))
        ok = True
        for step in range(4):
            X = orc.init_x(M, K, 100 + step)                  # every rank can compute the expected X ...
            def compute(Xd, Yd):
                Yd.copy_(torch.from_numpy(orc.base_tcsc(Xd.numpy(), t, b[lo:hi])))
            Yh = hc.step(compute, X if rank == 0 else None)   # ... but only rank 0 supplies it
            want = orc.base_tcsc(X, orc.tcsc(W), b)[:, lo:hi]
            ok = ok and np.array_equal(Yh.numpy(), want)
        dist.barrier()
        hc.close()
        open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,M", [(2, 6), (3, 6), (2, 7), (3, 8)])
def test_host_sharded_call_over_gloo(tmp_path, world, M):
    """Host X on rank 0 -> every rank uploads only its row block, the blocks are all-gathered, every
    rank computes its column shard: even and ragged row blocks."""
    mp.spawn(_worker_hostcall, args=(world, _free_port(), M, 48, 30, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"rank{r}.ok" for r in range(world)]


def test_bind_host_to_gpu_never_raises(tsg):
    """shard.bind_host_to_gpu is best effort: without NVML / a GPU it reports why nothing changed."""
    import os
    from ternary_spgemm_b200 import shard
    before = os.sched_getaffinity(0)
    msg = shard.bind_host_to_gpu(0)
    assert isinstance(msg, str) and msg
    assert os.sched_getaffinity(0) <= before          # never widens the affinity mask
    os.sched_setaffinity(0, before)
    os.environ["TSG_NO_NUMA_BIND"] = "1"
    try:
        assert shard.bind_host_to_gpu(0).startswith("off")
    finally:
        del os.environ["TSG_NO_NUMA_BIND"]
