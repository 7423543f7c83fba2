"""Multi-GPU driver: one process per GPU, zones sharded in contiguous ranges, one all-gather.

Replaces the MPI column distribution of parall.F90: rank p owns the zone range of parallPartion
(parall.F90:176-177, unit speeds), i.e. a contiguous slab of the zone-permuted state, plus the
observations within one search radius of its zones (the halo).  Observation-space arrays are
replicated in the reference (rrsqrt.F90:340-352 leaves Hxf/HSf/yo/R global); here each rank keeps only
its halo subset, in increasing global observation number, so local index sets map back one-to-one.
The analysed slabs are reassembled with a single NCCL all-gather (no other data-path collective).

The compute step is injected (`analyse`), so the same plumbing runs under gloo on CPUs in the tests
with the oracle as a stand-in, and under NCCL on B200s with Handle.local_analysis_dev.
"""
import numpy as np

from .api import LOC_HORIZONTAL, METRIC_CARTESIAN, WEIGHT_GASPARI_COHN, WEIGHT_GAUSSIAN

EARTH_RADIUS = 6378137.0


def partition(nzones, nranks):
    """first[p] .. first[p+1]: zones of rank p — parall.F90:176-177 with unit speeds (pure host twin of
    oakb200_partition_zones, so that planning needs no device)."""
    return np.array([(nzones * p) // nranks for p in range(nranks + 1)], dtype=np.int64)


def search_radius(corrLen, maxLen, weightfun):
    if weightfun == WEIGHT_GAUSSIAN:
        return np.asarray(maxLen, dtype=np.float64)
    if weightfun == WEIGHT_GASPARI_COHN:
        return 2.0 * np.asarray(corrLen, dtype=np.float64)
    return np.full(np.shape(corrLen), np.inf)


def halo_observations(zx, zy, radius, obs_x, obs_y, loctype=LOC_HORIZONTAL, metrictype=METRIC_CARTESIAN):
    """Global (0-based, increasing) numbers of the observations that can be relevant to any of the given
    zones: a conservative superset (bounding box of the zones grown by the largest radius, with slack);
    the exact predicate runs on the device.  Non-Cartesian / non-horizontal set-ups keep everything."""
    m = len(obs_x)
    if len(zx) == 0:
        return np.zeros(0, dtype=np.int64)
    rmax = float(np.max(radius)) if len(radius) else 0.0
    if loctype != LOC_HORIZONTAL or metrictype != METRIC_CARTESIAN or not np.isfinite(rmax):
        return np.arange(m, dtype=np.int64)
    pad = rmax * (1 + 1e-9) + 1e-9 * (np.abs(zx).max() + np.abs(zy).max() + 1.0)
    keep = ((obs_x >= zx.min() - pad) & (obs_x <= zx.max() + pad) & (obs_y >= zy.min() - pad) &
            (obs_y <= zy.max() + pad))
    return np.nonzero(keep)[0].astype(np.int64)


def phase_ranges(nzones, world, nphase):
    """Zone ranges for a pipelined run: the zones are cut in `nphase` consecutive phases, each phase is
    split over the ranks with the parallPartion formula.  first[j][p] .. first[j][p+1] = zones of rank p
    in phase j.  With nphase = 1 this is exactly parall.F90:176-177.  More phases let the all-gather of
    phase j overlap the analysis of phase j+1; the phases shrink towards the end (the gather of the last
    one is the only exposed communication) and hold a multiple of `world` zones so that slabs are equal."""
    if nphase <= 1:
        return [partition(nzones, world)]
    w = np.array([nphase + 1.0 - 0.5 * j for j in range(nphase)])   # e.g. 4 phases: 5 : 4.5 : 4 : 3.5 ... tapered
    w[-1] *= 0.5
    cut = np.concatenate([[0.0], np.cumsum(w)]) / w.sum()
    pb = [(int(round(c * nzones)) // world) * world for c in cut]
    pb[0], pb[-1] = 0, nzones
    pb = sorted(set(pb))
    return [pb[j] + partition(int(pb[j + 1] - pb[j]), world) for j in range(len(pb) - 1)]


class ShardPlan:
    """What rank `rank` of `world` owns (optionally inside one phase: `first` = that phase's boundaries)."""

    def __init__(self, zoneSize, zx, zy, corrLen, maxLen, obs_x, obs_y, rank, world, loctype=LOC_HORIZONTAL,
                 metrictype=METRIC_CARTESIAN, weightfun=WEIGHT_GAUSSIAN, first=None):
        zoneSize = np.asarray(zoneSize, dtype=np.int64)
        nz = zoneSize.size
        if first is None:
            first = partition(nz, world)
        first = np.asarray(first, dtype=np.int64)
        self.rank, self.world = rank, world
        self.first = first
        self.z0, self.z1 = int(first[rank]), int(first[rank + 1])
        start = np.concatenate([[0], np.cumsum(zoneSize)])
        self.row_first = start[first]                      # per-rank first rows
        self.r0, self.r1 = int(start[self.z0]), int(start[self.z1])
        self.n = int(start[-1])
        cl = np.broadcast_to(np.asarray(corrLen, dtype=np.float64), (nz,))
        ml = np.broadcast_to(np.asarray(maxLen, dtype=np.float64), (nz,))
        sl = slice(self.z0, self.z1)
        self.zoneSize = zoneSize[sl].astype(np.int32)
        self.zx = np.ascontiguousarray(zx[sl])
        self.zy = None if zy is None else np.ascontiguousarray(zy[sl])
        self.corrLen = np.ascontiguousarray(cl[sl])
        self.maxLen = np.ascontiguousarray(ml[sl])
        rad = search_radius(self.corrLen, self.maxLen, weightfun)
        self.obs_idx = halo_observations(self.zx, self.zy if self.zy is not None else np.zeros_like(self.zx), rad,
                                         np.asarray(obs_x), np.asarray(obs_y) if obs_y is not None else
                                         np.zeros_like(obs_x), loctype, metrictype)
        self.equal_slabs = bool(np.all(np.diff(self.row_first) == (self.r1 - self.r0)))


def allgather_slabs(dist, Sa_local, plan, out=None, wait=True):
    """Reassembles the member-major analysed slabs (N, n_loc) of all ranks into (N, n).

    Equal slabs: one all_gather_into_tensor per member row, issued asynchronously (a grouped collective
    on NCCL); with wait=False the work handles are returned so that the gather of one phase overlaps the
    analysis of the next.  Unequal slabs: all_gather on padded buffers (always waited)."""
    import torch
    N = Sa_local.shape[0]
    if out is None:
        out = torch.empty((N, plan.n), dtype=Sa_local.dtype, device=Sa_local.device)
    a, b = int(plan.row_first[0]), int(plan.row_first[-1])   # rows covered by this plan (a phase or everything)
    if plan.world == 1:
        out[:, a:b].copy_(Sa_local)
        return out if wait else []
    if plan.equal_slabs:
        # one grouped NCCL operation: the N per-member all-gathers (each lands in the contiguous rows
        # [a,b) of member k of the column-major n x N result) are coalesced into a single launch
        outs = [out[k, a:b] for k in range(N)]
        ins = [Sa_local[k] for k in range(N)]
        works = None
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
            with _coalescing_manager(async_ops=True) as cm:
                for o, i in zip(outs, ins):
                    dist.all_gather_into_tensor(o, i)
            works = [cm]
        except Exception:
            works = None
        if works is None:   # backend without coalescing support
            works = [dist.all_gather_into_tensor(o, i, async_op=True) for o, i in zip(outs, ins)]
        if not wait:
            return works
        for w in works:
            w.wait()
        return out
    sizes = np.diff(plan.row_first).astype(np.int64)
    nmax = int(sizes.max())
    buf = torch.zeros((N, nmax), dtype=Sa_local.dtype, device=Sa_local.device)
    buf[:, :Sa_local.shape[1]] = Sa_local
    parts = [torch.empty_like(buf) for _ in range(plan.world)]
    dist.all_gather(parts, buf)
    for p in range(plan.world):
        a, b = int(plan.row_first[p]), int(plan.row_first[p + 1])
        out[:, a:b] = parts[p][:, :b - a]
    return out if wait else []


class PeerResult:
    """The analysed state of the whole domain on every rank, filled by the ranks' own kernels (fused
    all-gather): each rank allocates Sa (N, n) member-major + xa (n) through the library (cudaMalloc + CUDA IPC
    handle), the handles are exchanged once with all_gather_object, and every rank maps the arrays of all the
    others (NVLink peer access).  `destinations(ld, row0)` is what Handle.set_peer_outputs takes: the apply
    kernel then stores this rank's rows into all `world` arrays, its own included.  Replaces parallGather
    (parall.F90:507-566) without a data-path collective; readers are ordered after the writers of all ranks by
    `fence()` (a one-element all-reduce enqueued behind the analysis)."""

    def __init__(self, dist, handle, N, n, rank, world, device):
        import torch
        self.dist, self.handle, self.rank, self.world = dist, handle, rank, world
        self.N, self.n = int(N), int(n)
        nel = self.N * self.n + self.n
        self.base, hd, self.opened, self.bases = None, None, [], []
        try:
            import os
            if os.environ.get("OAK_B200_TEST_PEER_FAIL") == str(rank):   # exercises the collective fall-back
                raise RuntimeError("forced")
            self.base, hd = handle.ipc_alloc(8 * nel)
        except Exception:
            hd = None
        handles = [None] * world
        dist.all_gather_object(handles, hd)        # every rank takes part, also one whose allocation failed
        if any(x is None for x in handles):
            self.free()
            raise RuntimeError("a rank could not allocate its result array")
        self.failed = None       # set instead of raising once peers may have mapped this rank's array:
        try:                      # the caller then tears everything down collectively (close_mappings, barrier, free)
            for r in range(world):
                if r == rank:
                    self.bases.append(self.base)
                else:
                    ptr = handle.ipc_open(handles[r])
                    self.opened.append(ptr)
                    self.bases.append(ptr)
        except Exception as e:
            self.failed = str(e)[:80]
            self.Sa = self.xa = None
            return

        class _Raw:   # __cuda_array_interface__ view of the library's allocation
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        flat = torch.as_tensor(_Raw(self.base, nel), device=device)
        self.Sa = flat[:self.N * self.n].view(self.N, self.n)
        self.xa = flat[self.N * self.n:]
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def destinations(self):
        Sa_ptrs = list(self.bases)
        xa_ptrs = [b + 8 * self.N * self.n for b in self.bases]
        return Sa_ptrs, xa_ptrs

    def fence(self):
        """Stream-ordered barrier over the ranks: returns (on the stream) once every rank's analysis kernels,
        hence their stores into this rank's arrays, are complete."""
        self.dist.all_reduce(self._flag)

    def close_mappings(self):
        for ptr in self.opened:
            self.handle.ipc_close(ptr)
        self.opened = []

    def free(self):
        if self.base:
            self.handle.ipc_free(self.base)
            self.base = None

    def close(self):
        """Collective: unmap the peers' arrays, wait until nobody maps this rank's any more, free it."""
        self.close_mappings()
        self.dist.barrier()
        self.free()


class MulticastResult:
    """The analysed state of the whole domain on every rank, filled through NVSwitch MULTICAST: each rank creates its
    result array (Sa (N, n) member-major + xa (n)) as a physical allocation of its own (cuMemCreate), all ranks bind
    their arrays at offset 0 of ONE multicast object (cuMulticastCreate on rank 0, shared as a POSIX file descriptor
    over a Unix socket; cuMulticastAddDevice; cuMulticastBindMem) and map the object's multicast address.  A store to
    the multicast address (multimem.st, k_push_mc) is replicated by the switch into every rank's array, so a rank sends
    its slab ONCE instead of once per peer (8 GPUs, C3: 1.9 GB instead of 13.4 GB of NVLink egress per rank and step).
    Same interface as PeerResult (Sa, xa, fence, close); `multicast_pointers()` is what Handle.set_multicast_output
    takes.  Raises if the device or driver has no multicast support (the caller falls back to PeerResult / NCCL).
    Replaces parallGather (parall.F90:507-566)."""

    def __init__(self, dist, N, n, rank, world, device, sock_dir="/tmp"):
        import os
        import socket
        import torch
        from cuda.bindings import driver as drv
        self.drv, self.dist, self.rank, self.world = drv, dist, rank, world
        self.N, self.n = int(N), int(n)
        self._maps = []

        def ck(r):
            if r[0] != drv.CUresult.CUDA_SUCCESS:
                raise RuntimeError("CUDA driver: " + str(r[0]))
            return r[1] if len(r) == 2 else (r[1:] if len(r) > 2 else None)
        self._ck = ck
        ck(drv.cuInit(0))
        dev = ck(drv.cuDeviceGet(device.index))
        if not ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)):
            raise RuntimeError("device has no multicast support")
        fdtype = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        nbytes = 8 * (self.N * self.n + self.n)
        mcprop = drv.CUmulticastObjectProp()
        mcprop.numDevices = world
        mcprop.handleTypes = fdtype
        mcprop.flags = 0
        mcprop.size = nbytes
        gran = ck(drv.cuMulticastGetGranularity(mcprop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_MINIMUM))
        aprop = drv.CUmemAllocationProp()
        aprop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        aprop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        aprop.location.id = device.index
        aprop.requestedHandleTypes = fdtype
        g2 = ck(drv.cuMemGetAllocationGranularity(aprop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
        gran = max(int(gran), int(g2))
        size = (nbytes + gran - 1) // gran * gran
        mcprop.size = size
        self.size = size
        # this rank's physical array and its ordinary (unicast) mapping: what the readers use
        self.phys = ck(drv.cuMemCreate(size, aprop, 0))
        acc = drv.CUmemAccessDesc()
        acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = device.index
        acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        self.va = ck(drv.cuMemAddressReserve(size, gran, 0, 0))
        ck(drv.cuMemMap(self.va, size, 0, self.phys, 0))
        ck(drv.cuMemSetAccess(self.va, size, [acc], 1))
        # the multicast object: created by rank 0, imported by the others from its file descriptor
        path = os.path.join(sock_dir, "oak_b200_mc_%s.sock" % os.environ.get("MASTER_PORT", "0"))
        if rank == 0:
            self.mc = ck(drv.cuMulticastCreate(mcprop))
            fd = int(ck(drv.cuMemExportToShareableHandle(self.mc, fdtype, 0)))
            if os.path.exists(path):
                os.unlink(path)
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            srv.listen(world)
            dist.barrier()                                   # the socket exists
            for _ in range(world - 1):
                c, _a = srv.accept()
                socket.send_fds(c, [b"mc"], [fd])
                c.close()
            srv.close()
            os.close(fd)
            os.unlink(path)
        else:
            dist.barrier()
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path)
            _msg, fds, _f, _a = socket.recv_fds(c, 16, 1)
            c.close()
            self.mc = ck(drv.cuMemImportFromShareableHandle(fds[0], fdtype))
            os.close(fds[0])
        ck(drv.cuMulticastAddDevice(self.mc, dev))
        dist.barrier()                                       # every device has been added
        ck(drv.cuMulticastBindMem(self.mc, 0, self.phys, 0, size, 0))
        self.mcva = ck(drv.cuMemAddressReserve(size, gran, 0, 0))
        ck(drv.cuMemMap(self.mcva, size, 0, self.mc, 0))
        ck(drv.cuMemSetAccess(self.mcva, size, [acc], 1))
        dist.barrier()                                       # every rank has bound and mapped

        class _Raw:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        flat = torch.as_tensor(_Raw(int(self.va), self.N * self.n + self.n), device=device)
        self.Sa = flat[:self.N * self.n].view(self.N, self.n)
        self.xa = flat[self.N * self.n:]
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def multicast_pointers(self):
        return int(self.mcva), int(self.mcva) + 8 * self.N * self.n

    def fence(self):
        self.dist.all_reduce(self._flag)

    def close(self):
        """Collective."""
        import torch
        drv, ck = self.drv, self._ck
        self.Sa = self.xa = None
        torch.cuda.synchronize()
        self.dist.barrier()
        try:
            ck(drv.cuMemUnmap(self.mcva, self.size)); ck(drv.cuMemAddressFree(self.mcva, self.size))
            ck(drv.cuMemUnmap(self.va, self.size)); ck(drv.cuMemAddressFree(self.va, self.size))
            self.dist.barrier()
            ck(drv.cuMemRelease(self.mc)); ck(drv.cuMemRelease(self.phys))
        except Exception:
            pass
