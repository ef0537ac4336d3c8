"""x-strip decomposition of a structured SEM2DPACK box across the GPUs of one node.

One process per GPU.  Rank r owns element columns [x_lo(r), x_hi(r)) of the global box as its own
`CartEngine` (built with `halo_left` / `halo_right`), so the GLL lattice column at each interface
exists on both neighbours.  After every force evaluation the engine packs its partial sums of those
columns, calls the exchange hook registered here, and adds the neighbour's partial sums
(own + neighbour's: the same two addends on both sides, hence bit-identical values).  The hook runs on
the engine's side stream, so the transfer overlaps the interior strips that are still computing.
There is no other collective on the data path: sources, receivers, the fault and the absorbing
sides are node-local and replicated on interface nodes.

`exchange_halos` is the whole protocol (one batched send/recv pair per neighbour); it works on any
torch.distributed backend, which is how the CPU tests drive it over gloo.
"""
import ctypes as C
import threading

import torch
import torch.distributed as dist

from . import capi


def partition(nx_global, world):
    """Element-column range [lo, hi) of every rank: as even as possible, wider strips first."""
    base, rem = divmod(nx_global, world)
    out, lo = [], 0
    for r in range(world):
        w = base + (1 if r < rem else 0)
        out.append((lo, lo + w))
        lo += w
    return out


def exchange_halos(send_left, send_right, recv_left, recv_right, rank, world, group=None):
    """Neighbour exchange of the interface partial sums.  A tensor is None on a side without a
    neighbour.  Returns the list of outstanding requests (already issued as one batch)."""
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, send_left, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_left, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, send_right, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_right, rank + 1, group))
    return dist.batch_isend_irecv(ops) if ops else []


class _DevView:
    """Lets torch adopt a raw device pointer owned by the engine (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def halo_tensors(engine, precision=8):
    """(send_left, send_right, recv_left, recv_right) torch views of the engine's halo staging buffers."""
    L = engine.L
    count = C.c_int64()
    send = (C.c_void_p * 2)()
    recv = (C.c_void_p * 2)()
    engine._ck(L.s2d_halo_info(engine.h, C.byref(count), send, recv))
    typestr = "<f8" if precision == 8 else "<f4"
    dev = torch.device("cuda", torch.cuda.current_device())

    def view(p):
        return torch.as_tensor(_DevView(p, count.value, typestr), device=dev) if p else None
    return view(send[0]), view(send[1]), view(recv[0]), view(recv[1])


_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


def attach_halo_exchange(engine, rank, world, group=None, precision=8):
    """Registers the NCCL neighbour exchange as the engine's halo hook (s2d_halo_set_exchange)."""
    sl, sr, rl, rr = halo_tensors(engine, precision)

    def hook(_user, stream_ptr):
        try:
            ext = torch.cuda.ExternalStream(stream_ptr)
            with torch.cuda.stream(ext):
                for req in exchange_halos(sl, sr, rl, rr, rank, world, group):
                    req.wait()  # orders the side stream after the transfer; does not block the host
            return 0
        except Exception as ex:  # surfaces as S2D_ESTATE from s2d_step
            print(f"[sem2dpack_b200.strips] halo exchange failed on rank {rank}: {ex}", flush=True)
            return 1

    cb = _HOOK(hook)
    engine._halo_cb = (cb, sl, sr, rl, rr)  # keep alive
    engine._ck(engine.L.s2d_halo_set_exchange(engine.h, C.cast(cb, C.c_void_p), None))


def attach_peer_exchange(engine, rank, world, group=None):
    """Direct peer-memory halo exchange over NVLink (s2d_halo_ipc_export / s2d_halo_ipc_open): every
    rank publishes CUDA IPC handles of its receive slots and flags, the neighbours map them, and from
    then on the engine's own kernels write the interface partial sums into the neighbour's memory
    and signal through device flags -- no host hook and no NCCL call on the step path.
    torch.distributed only carries the 192-byte handles once."""
    blob = (C.c_ubyte * capi.HALO_IPC_BYTES)()
    engine._ck(engine.L.s2d_halo_ipc_export(engine.h, blob))
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes(blob), group=group)

    def buf(r):
        return (C.c_ubyte * capi.HALO_IPC_BYTES).from_buffer_copy(blobs[r]) if 0 <= r < world else None
    left, right = buf(rank - 1), buf(rank + 1)
    engine._peer_blobs = (left, right)
    engine._ck(engine.L.s2d_halo_ipc_open(engine.h, left, right))


class LocalStrips:
    """Several x-strips of one box on ONE GPU inside one process, stepped by one thread each and
    exchanging through device copies -- the same engine-side halo path as the multi-GPU run, used by
    the single-GPU parity tests (whole box vs strips must agree bit for bit)."""

    def __init__(self, engines, precision=8, peer=False):
        self.engines = engines
        self.n = len(engines)
        self.barrier = threading.Barrier(self.n)
        if peer:  # the device-side protocol of the multi-GPU run: neighbours' slots and flags, no hook
            ptrs = []
            for e in engines:
                recv = (C.c_void_p * 2)()
                flags = C.c_void_p()
                e._ck(e.L.s2d_halo_peer_buffers(e.h, recv, C.byref(flags)))
                ptrs.append((recv[0], recv[1], flags.value))
            for r, e in enumerate(engines):
                lr = lf = rr = rf = None
                if r > 0:
                    lr, lf = ptrs[r - 1][1], ptrs[r - 1][2] + 8
                if r < self.n - 1:
                    rr, rf = ptrs[r + 1][0], ptrs[r + 1][2]
                e._ck(e.L.s2d_halo_set_peers(e.h, lr, rr, lf, rf))
            return
        self.bufs = [halo_tensors(e, precision) for e in engines]
        self._cbs = []
        for r, e in enumerate(engines):
            cb = _HOOK(self._make_hook(r))
            self._cbs.append(cb)
            e._ck(e.L.s2d_halo_set_exchange(e.h, C.cast(cb, C.c_void_p), None))

    def _make_hook(self, r):
        def hook(_user, stream_ptr):
            try:
                ext = torch.cuda.ExternalStream(stream_ptr)
                ext.synchronize()            # my partial sums are packed
                self.barrier.wait()          # ... and so are everybody's
                with torch.cuda.stream(ext):
                    if r > 0:
                        self.bufs[r][2].copy_(self.bufs[r - 1][1])       # my recv_left <- left's send_right
                    if r < self.n - 1:
                        self.bufs[r][3].copy_(self.bufs[r + 1][0])       # my recv_right <- right's send_left
                ext.synchronize()
                self.barrier.wait()          # nobody repacks before everybody has copied
                return 0
            except Exception as ex:
                print(f"[sem2dpack_b200.strips] local exchange failed on strip {r}: {ex}", flush=True)
                self.barrier.abort()
                return 1
        return hook

    def run(self, fn):
        """Calls fn(rank, engine) on one thread per strip and returns the results in rank order."""
        out, err = [None] * self.n, [None] * self.n

        def work(r):
            try:
                out[r] = fn(r, self.engines[r])
            except Exception as ex:  # noqa: BLE001
                err[r] = ex
                self.barrier.abort()
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for ex in err:
            if ex is not None:
                raise ex
        return out
