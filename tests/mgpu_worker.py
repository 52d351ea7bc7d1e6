"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun, NCCL).  The N-sharded CUDA path —
tsg_tcsc_from_dense_cols + tsg_spmm_dev / tsg_spmm + gather_columns — against the oracle's BaseTCSC
on the whole matrix, for every way X reaches the ranks (NCCL broadcast, peer reads of rank 0's
symmetric memory, host shared memory)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle.pyoracle import Oracle

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tsg = ge.load_package()
    from ternary_spgemm_b200 import shard
    orc = Oracle()
    done = []
    try:
        for (M, K, N, s) in ((1, 1024, 3000, 4), (40, 1024, 3000, 4), (300, 2048, 4100, 8)):
            W = orc.generate_sparse_matrix(K, N, s, 77)           # same on every rank
            o_full = orc.tcsc(W)
            lo, hi = tsg.shard_columns(N, world, rank)
            t = tsg.TCSC(W, col_range=(lo, hi))                   # tsg_tcsc_from_dense_cols
            o_loc = orc.tcsc(np.ascontiguousarray(W[:, lo:hi]))
            for got, exp in zip(t.export(), o_loc.arrays):
                assert np.array_equal(got, exp), "shard arrays differ from TCSC(W[:, lo:hi])"
            rng = np.random.default_rng(5)                        # same stream on every rank
            b = rng.uniform(-1, 1, N).astype(np.float32)
            al = rng.uniform(0.01, 0.3, N).astype(np.float32)
            db, da = torch.from_numpy(b[lo:hi].copy()).to(dev), torch.from_numpy(al[lo:hi].copy()).to(dev)
            for regime in ("int", "real"):
                X = (orc.init_x(M, K, 9) if regime == "int" else rng.uniform(-1, 1, (M, K)).astype(np.float32))
                want = orc.base_tcsc_prelu(X, o_full, b, al)
                Xsrc = X if rank == 0 else np.zeros_like(X)       # only rank 0 holds the batch

                def check(Yloc, how):
                    Y = shard.gather_columns(Yloc, N).cpu().numpy()
                    if regime == "int":
                        assert np.array_equal(Y, want), (how, M, regime)
                    else:
                        err = np.abs(Y.astype(np.float64) - want).max() / np.abs(want).max()
                        assert err <= 1e-5, (how, M, regime, err)
                    done.append(how)

                # (1) NCCL broadcast of X, device-pointer call
                dX = torch.from_numpy(Xsrc).to(dev)
                shard.broadcast_x(dX, src=0)
                dY = torch.empty(M, hi - lo, device=dev)
                t.spmm_dev(dX, db, dY, M, alpha=da, stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                check(dY, "broadcast_x")
                # (2) peer reads of rank 0's symmetric-memory X inside the kernel
                try:
                    px = shard.PeerX(M, K, dev)
                    for step in range(3):                          # both buffers, reuse
                        xin = px.stage(torch.from_numpy(Xsrc).to(dev))
                        dY.zero_()
                        t.spmm_dev(xin, db, dY, M, alpha=da, stream=torch.cuda.current_stream().cuda_stream)
                        torch.cuda.synchronize()
                        check(dY, "PeerX")
                except (RuntimeError, AttributeError, ImportError) as e:   # symmetric memory unavailable on this box
                    if rank == 0:
                        print(f"PeerX skipped: {type(e).__name__}: {str(e)[:100]}", flush=True)
                # (3) X through host shared memory, host-pointer call on every rank
                hx = shard.HostSharedX(M, K)
                for step in range(3):
                    x = hx.next(X if rank == 0 else None)
                    Yh = t.spmm(x, b[lo:hi], al[lo:hi])
                    hx.done()
                    check(torch.from_numpy(Yh).to(dev), "HostSharedX")
                dist.barrier()
                hx.close()
                # (4) host X, 1/G of it uploaded per rank, NVLink all-gather, Y slices back to the host
                hc = shard.HostShardedCall(M, K, hi - lo, dev)
                for step in range(3):
                    Yh = hc.step(lambda Xd, Yd: t.spmm_dev(Xd, db, Yd, M, alpha=da,
                                                           stream=torch.cuda.current_stream().cuda_stream),
                                 X if rank == 0 else None)
                    check(Yh.to(dev), "HostShardedCall")
                dist.barrier()
                hc.close()
            t.close()
        if rank == 0:
            from collections import Counter
            print("mgpu ok", world, dict(Counter(done)), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
