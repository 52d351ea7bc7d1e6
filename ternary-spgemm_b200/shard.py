"""N-column sharding across the GPUs of one box (north-star subsystem 4).

Column n of Y depends only on column n of W (reference cpp_impl/comp.h:39-63), and TCSC is
column-major, so rank r of G owns the contiguous column slice [N·r/G, N·(r+1)/G): a contiguous
range of all four TCSC arrays.  Per input there is exactly ONE collective — the broadcast of X
from rank 0 (M·K floats over NVLink/NVSwitch via NCCL) — and no reduction: every rank writes its
own Y[:, lo:hi].  `gather_columns` exists for verification / callers that want Y assembled; it
is not on the timed path.

torch.distributed is plumbing here (process group, NCCL/gloo); the compute is whatever
`compute(X, lo, hi)` the caller passes — libtsg's kernels on the GPU box, the checker in the
CPU gloo tests.
"""
from __future__ import annotations

from . import shard_columns  # noqa: F401  (re-export: the partition rule lives with the format)


def broadcast_x(X, src: int = 0, group=None):
    """Replicate the activation batch: the only data-path collective."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(X, src=src, group=group)
    return X


class PeerX:
    """X replicated WITHOUT a broadcast: rank `src` owns the activation buffer in symmetric memory
    (NVLink / NVSwitch peer mapping) and every other rank's SpMM kernel reads it in place through
    its peer pointer — the transfer happens inside the consuming kernel (for the tensor-core path
    it is the one read of X by its split kernel; for the decode kernels a few KB per CTA), so there
    is no collective launch and no second copy of X in HBM.

    Two buffers alternate, so one device-side barrier per step is enough: a rank enqueues barrier i
    after its kernel of step i-1, hence when barrier i has completed everywhere, buffer (i-1)%2 is
    free again for the owner's write of step i+1.

        px = PeerX(M, K, device)                 # collective: all ranks
        x = px.stage(host_or_device_X)           # rank src copies, everyone barriers; returns the view
        matrix.spmm_dev(x, b, Y, M, ...)         # kernel pulls X over NVLink
    """

    def __init__(self, M: int, K: int, device, *, group=None, src: int = 0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.src, self.rank = src, dist.get_rank(group)
        g = group if group is not None else dist.group.WORLD
        self.local, self.hdl, self.view = [], [], []
        for _ in range(2):
            t = symm.empty((M, K), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, g)
            self.local.append(t)
            self.hdl.append(h)
            self.view.append(t if self.rank == src else h.get_buffer(src, (M, K), torch.float32))
        self.step = 0

    def stage(self, X):
        """Publish this step's X (only rank `src` reads its argument).  Enqueued on the current
        stream; returns the tensor every rank passes to its kernel."""
        i = self.step & 1
        self.step += 1
        if self.rank == self.src:
            self.local[i].copy_(X, non_blocking=True)
        self.hdl[i].barrier(channel=0)
        return self.view[i]


class HostSharedX:
    """X replicated through HOST shared memory, for callers whose X lives on the host (the
    reference's comp_func contract): rank `src` publishes each step's X in a POSIX shared-memory
    block every rank has mapped (registered with CUDA when a GPU is present), and every rank's
    host-pointer call pulls X over ITS OWN PCIe link — G links in parallel, no inter-GPU traffic, no
    device-side barrier, no collective.  Two buffers alternate; a sequence word orders the steps and
    per-rank acknowledgement words keep the publisher from overwriting a buffer that is still read.

        hx = HostSharedX(M, K)                  # collective: all ranks
        x = hx.next(X_host)                     # src copies X in and publishes; the others wait for it
        matrix.spmm_host_ptr(x.ctypes.data, ...)# or matrix.spmm(x, b)
        hx.done()                               # this rank has consumed the step
    """

    def __init__(self, M: int, K: int, *, group=None, src: int = 0):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory

        import os
        self.src, self.rank = src, dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.M, self.K = M, K
        self._hdr = (64 * (1 + self.world) + 4095) // 4096 * 4096   # one cache line per counter; X page-aligned
        self._xbytes = (M * K * 4 + 4095) // 4096 * 4096
        self.timeout_s = float(os.environ.get("TSG_SHM_TIMEOUT_S", "60"))
        size = self._hdr + 2 * self._xbytes
        # Collective and failure-proof: the creator reports success or the reason, every rank
        # learns it, and a rank that cannot attach makes ALL ranks raise (callers fall back together).
        import torch

        payload = [None, None]
        if self.rank == src:
            try:
                st = os.statvfs("/dev/shm")
                free = st.f_bavail * st.f_frsize
                if free < size + (16 << 20):                 # writing past a full tmpfs is a SIGBUS, not an error
                    raise OSError(f"/dev/shm has {free >> 20} MiB free, {size >> 20} MiB needed")
                self._shm = shared_memory.SharedMemory(create=True, size=size)
                self._shm.buf[: self._hdr] = bytes(self._hdr)
                payload[0] = self._shm.name
            except Exception as e:
                payload[1] = f"{type(e).__name__}: {e}"
        dist.broadcast_object_list(payload, src=src, group=group)
        if payload[0] is None:
            raise RuntimeError("host shared memory unavailable on the publishing rank: " + str(payload[1]))
        attached = 1
        if self.rank != src:
            try:
                self._shm = shared_memory.SharedMemory(name=payload[0])
                try:                                         # the creator unlinks; do not let this process's
                    from multiprocessing import resource_tracker  # tracker do it a second time at exit
                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
            except Exception:
                attached = 0
        on_gpu = dist.get_backend(group) == "nccl"
        flag = torch.tensor([attached], dtype=torch.int32, device="cuda" if on_gpu else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if self.rank == src:
                self._shm.close()
                self._shm.unlink()
            elif attached:
                self._shm.close()
            raise RuntimeError("host shared memory could not be attached on every rank")
        self._cnt = np.ndarray((1 + self.world, 8), dtype=np.int64, buffer=self._shm.buf)  # [0]=seq, [1+r]=ack of r
        self._x = [np.ndarray((M, K), dtype=np.float32, buffer=self._shm.buf, offset=self._hdr + i * self._xbytes)
                   for i in range(2)]
        # counters are published / consumed through release / acquire accesses (libtsg's two host
        # helpers: plain numpy stores order nothing on weakly ordered hosts such as Grace)
        from . import lib
        L = lib()
        base = self._cnt.ctypes.data
        self._seq_addr = base
        self._ack_addr = [base + 64 * (1 + r) for r in range(self.world)]
        self._store, self._load = L.tsg_host_store_release_i64, L.tsg_host_load_acquire_i64
        # pinned for this rank's DMA (large calls); a failure is reported, the copies then run from
        # pageable memory (slower, still correct)
        self._registered, self.register_error = False, None
        if torch.cuda.is_available():
            rc = int(torch.cuda.cudart().cudaHostRegister(self._x[0].ctypes.data, 2 * self._xbytes, 0))
            self._registered = rc == 0
            if rc != 0:
                self.register_error = f"cudaHostRegister failed with status {rc}"
                import warnings
                warnings.warn("HostSharedX: " + self.register_error + "; X is copied from pageable memory")
        self.step = 0
        dist.barrier(group=group)

    def _wait(self, ready, what):
        """Spin (then yield) until ready() holds; every rank gives up after timeout_s with an error
        instead of hanging on a rank that died or never called done()."""
        import time
        spins, t0 = 0, None
        while not ready():
            spins += 1
            if spins > 2000:
                time.sleep(0)
                if t0 is None:
                    t0 = time.monotonic()
                elif time.monotonic() - t0 > self.timeout_s:
                    raise TimeoutError(f"HostSharedX rank {self.rank}: waited {self.timeout_s:.0f} s for {what} "
                                       f"(step {self.step}); a rank died or did not call done()")

    def next(self, X=None):
        """Step forward: `src` copies X (numpy / CPU tensor, M×K fp32) into the free buffer and
        publishes it; every rank gets the published view back."""
        import numpy as np

        self.step += 1
        s, buf = self.step, self._x[self.step & 1]
        if self.rank == self.src:
            # buffer s%2 last carried step s-2: every rank must have acknowledged it
            self._wait(lambda: min(self._load(a) for a in self._ack_addr) >= s - 2, "the readers of the buffer")
            if X is not None:                                # None: the producer wrote `buffers()[s & 1]` in place
                buf[...] = np.asarray(X, dtype=np.float32).reshape(self.M, self.K)
            self._store(self._seq_addr, s)                   # release: the data is visible before the word
        else:
            self._wait(lambda: self._load(self._seq_addr) >= s, "the publisher")
        return buf

    def buffers(self):
        """The two M×K views steps alternate between (step s uses index s & 1): a producer that
        writes X here directly publishes with next(None) — no copy on the publishing side."""
        return self._x

    def done(self):
        self._store(self._ack_addr[self.rank], self.step)

    def close(self):
        try:
            if self._registered:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self._x[0].ctypes.data)
            self._cnt = None
            self._x = None
            self._shm.close()
            if self.rank == self.src:
                self._shm.unlink()
        except Exception:
            pass


def bind_host_to_gpu(device_index: int) -> str:
    """Pin the calling process to the CPUs NVML reports as local to this GPU (same socket / NUMA
    node as its PCIe root), so that the pinned host buffers it allocates afterwards are placed
    next to the link they travel over — one process per GPU otherwise lands wherever the launcher
    put it and eight Y slices may cross the inter-socket link.  Returns what was done (for the
    bench line); never raises: without NVML, or inside a cpuset that excludes those CPUs, nothing
    changes.  TSG_NO_NUMA_BIND=1 turns it off."""
    import os
    if os.environ.get("TSG_NO_NUMA_BIND"):
        return "off (TSG_NO_NUMA_BIND)"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        prop = torch.cuda.get_device_properties(device_index)
        try:
            bus = "%08x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return f"unchanged ({len(allowed)} usable cpus, {len(local & allowed)} of them local to the GPU)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} of {len(allowed)} usable cpus (local to GPU {device_index})"
    except Exception as e:  # noqa: BLE001 — a missing NVML or a refused affinity call must not stop a run
        return f"unchanged ({type(e).__name__}: {e})"


class HostShardedCall:
    """One sharded call with HOST operands — X lives in the host memory of rank `src` (the
    reference's comp_func contract, cpp_impl/common.h:12), every rank returns its Y[:, lo:hi] slice
    to host memory — with every PCIe link carrying only 1/G of X:

        publish   rank src puts X in host shared memory (HostSharedX: no copy when it writes it there)
        upload    rank r copies ROW BLOCK r of X (M/G rows) to its GPU over its own PCIe link
        exchange  one all-gather of the row blocks over NVLink/NVSwitch (NCCL) assembles X on every GPU
        compute   the rank's kernels on its column shard (tsg_spmm_dev)
        download  the rank's Y slice to pinned host memory

    Against every rank pulling all of X (HostSharedX + tsg_spmm): the H2D bytes per link fall from
    4·M·K to 4·M·K/G and the exchange runs at NVLink speed.  Needs M % G == 0 for the single-tensor
    all-gather (otherwise the blocks are gathered one by one)."""

    def __init__(self, M: int, K: int, n_local: int, device, *, group=None, src: int = 0):
        import torch
        import torch.distributed as dist

        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.M, self.K = M, K
        self.hx = HostSharedX(M, K, group=group, src=src)
        self.rows = [((M * r) // self.world, (M * (r + 1)) // self.world) for r in range(self.world)]
        self.even = M % self.world == 0 and dist.get_backend(group) == "nccl"
        self.Xd = torch.empty(M, K, device=device)
        self.Yd = torch.empty(M, n_local, device=device)
        self.Yh = torch.empty(M, n_local)
        if device.type == "cuda":
            self.Yh = self.Yh.pin_memory()

    def buffers(self):
        return self.hx.buffers()

    def step(self, compute, X=None):
        """compute(Xd, Yd) enqueues the rank's kernels on the current stream.  Returns the pinned
        host tensor holding this rank's Y slice (valid until the next step)."""
        import torch
        import torch.distributed as dist

        x = self.hx.next(X)
        lo, hi = self.rows[self.rank]
        mine = self.Xd[lo:hi]
        mine.copy_(torch.from_numpy(x[lo:hi]), non_blocking=True)      # 1/G of X over this rank's PCIe link
        if self.world > 1:
            if self.even:
                dist.all_gather_into_tensor(self.Xd, mine, group=self.group)
            else:
                dist.all_gather([self.Xd[a:b] for a, b in self.rows], mine, group=self.group) \
                    if len({b - a for a, b in self.rows}) == 1 else self._gather_ragged(mine)
        compute(self.Xd, self.Yd)
        self.Yh.copy_(self.Yd, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self.hx.done()
        return self.Yh

    def _gather_ragged(self, mine):
        import torch.distributed as dist
        for r, (a, b) in enumerate(self.rows):                         # ragged row blocks: one broadcast each
            if b > a:
                dist.broadcast(self.Xd[a:b], src=r, group=self.group)

    def close(self):
        self.hx.close()


def sharded_spmm(X, N: int, compute, *, group=None, src: int = 0):
    """Run one step of the sharded path on this rank: broadcast X, compute the local column
    slice.  Returns (Y_local, (lo, hi))."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_columns(N, world, rank)
    broadcast_x(X, src=src, group=group)
    return compute(X, lo, hi), (lo, hi)


def gather_columns(Y_local, N: int, group=None):
    """Assemble the full M×N result on every rank from the per-rank column slices (uneven
    slices allowed).  Verification helper — not part of the timed path."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return Y_local
    world = dist.get_world_size(group)
    M = Y_local.shape[0]
    widths = [shard_columns(N, world, r)[1] - shard_columns(N, world, r)[0] for r in range(world)]
    wmax = max(widths)
    pad = torch.zeros(M, wmax, dtype=Y_local.dtype, device=Y_local.device)
    pad[:, : Y_local.shape[1]] = Y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, :w] for p, w in zip(parts, widths)], dim=1)
