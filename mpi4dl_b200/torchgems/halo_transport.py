"""Halo-strip transports for the spatial layers.

Replaces the reference's tagged `dist.isend/irecv` on CUDA-aware MPI with a device-wide
`torch.cuda.synchronize()` before every message (spatial.py:336-403):

* PeerTransport (default on GPUs of one node): every rank owns one cudaMalloc'ed *mailbox*
  arena, mapped into its neighbours with CUDA IPC.  One pack kernel (spc_halo_pack) writes all
  outgoing strips straight into the neighbours' arenas over NVLink/NVSwitch (P2P stores), then a
  release-store at system scope publishes a per-slot sequence number; the receiver's stream
  waits on the flag with an acquire-load spin kernel.  No host synchronisation, no tags.
* DistTransport: the same strips through `torch.distributed.batch_isend_irecv` (NCCL on GPUs,
  gloo on CPU for the plumbing tests).  Selected with SPCONV_HALO_TRANSPORT=dist or
  automatically when peer mapping is not possible.

`exchange_strips()` is the device-agnostic communication core (used by DistTransport and by the
CPU/gloo tests of the neighbour arithmetic).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from .. import _lib

_DIRS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0), (1, 1)]


def strip_shape(i, N, Cc, H, W, hh, hw):
    """Shape of strip i (reference spatial.py:311-334 get_shapes_recv)."""
    dr, dc = _DIRS[i]
    return (N, Cc, H if dr == 0 else hh, W if dc == 0 else hw)


def exchange_strips(send, recv, ranks, group=None):
    """Send send[i] to ranks[i] and receive recv[i] from ranks[i] for every i with a neighbour.
    Strip i travels to the neighbour in direction i, who receives it as ITS strip 8-i (the
    reference pairs send tag[i] with recv tag[8-i], spatial.py:170-172).  Between any ordered
    pair of ranks there is at most one strip per exchange, so no tags are needed."""
    ops = []
    for i in range(9):
        if i == 4 or send[i] is None:
            continue
        ops.append(dist.P2POp(dist.isend, send[i], ranks[i], group))
        ops.append(dist.P2POp(dist.irecv, recv[i], ranks[i], group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pack(x, hh, hw, ptrs):
    L = _lib.lib()
    N, Cc, H, W = x.shape
    arr = (C.c_void_p * 9)(*[C.c_void_p(p) if p else C.c_void_p(None) for p in ptrs])
    _lib.check(L.spc_halo_pack(N, Cc, H, W, hh, hw, _lib.dtype_code(x.dtype), C.c_void_p(x.data_ptr()),
                               C.byref(arr), _stream()), "spc_halo_pack")


class DistTransport:
    name = "dist"

    def exchange(self, layer, x, hh, hw, mask, ranks):
        N, Cc, H, W = x.shape
        send = [None] * 9
        recv = [None] * 9
        for i in range(9):
            if i != 4 and mask[i]:
                shp = strip_shape(i, N, Cc, H, W, hh, hw)
                send[i] = torch.empty(shp, dtype=x.dtype, device=x.device)
                recv[i] = torch.empty(shp, dtype=x.dtype, device=x.device)
        _pack(x, hh, hw, [t.data_ptr() if t is not None else 0 for t in send])
        exchange_strips(send, recv, ranks)
        return recv


class _CudaMem:
    """Expose raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerTransport:
    """CUDA-IPC mailbox transport (see module docstring).  Layout of the arena is symmetric
    across the ranks of a spatial group: every rank executes the same layer sequence on
    equally-shaped tiles (the reference requires power-of-two image and part counts,
    train_spatial.py:33-58), so a layer's slot offset is the same everywhere and a sender can
    address its neighbour's slot without a handshake per layer."""

    name = "peer"
    FLAGS_PER_LAYER = 36  # arrival[2][9] + ack[2][9]

    def __init__(self, device):
        self.device = device
        L = _lib.lib()
        self.arena_bytes = int(os.environ.get("SPCONV_ARENA_MB", "256")) << 20
        self.nflags = 1 << 16
        mb = C.c_void_p()
        _lib.check(L.spc_mailbox_create(C.byref(mb), self.arena_bytes, self.nflags), "spc_mailbox_create")
        self.mb = mb
        self.base = L.spc_mailbox_data(mb)
        self.arena = torch.as_tensor(_CudaMem(self.base, self.arena_bytes), device=device)
        handle = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(L.spc_mailbox_export(mb, handle), "spc_mailbox_export")
        # the 64-byte IPC handle travels through torch.distributed once per neighbour
        self.hdev = device if (dist.is_initialized() and dist.get_backend() == "nccl") else torch.device("cpu")
        self.handle = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.hdev)
        self.peers = {}       # rank -> (mailbox ptr, data base ptr)
        self.data_top = 0
        self.flag_top = 0

    def _peer(self, rank):
        if rank not in self.peers:
            L = _lib.lib()
            theirs = torch.empty(_lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.hdev)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, self.handle, rank),
                                           dist.P2POp(dist.irecv, theirs, rank)])
            for r in reqs:
                r.wait()
            raw = bytes(theirs.cpu().numpy().tobytes())
            mb = C.c_void_p()
            _lib.check(L.spc_mailbox_open(C.byref(mb), raw, self.arena_bytes, self.nflags), "spc_mailbox_open")
            self.peers[rank] = (mb, L.spc_mailbox_data(mb))
        return self.peers[rank]

    def _slot(self, layer, x, hh, hw):
        key = (tuple(x.shape), x.dtype, hh, hw)
        slots = layer.__dict__.setdefault("_halo_slots", {})
        if key not in slots:
            N, Cc, H, W = x.shape
            off, offs = 0, []
            for i in range(9):
                offs.append(off)
                if i != 4:
                    n = 1
                    for s in strip_shape(i, N, Cc, H, W, hh, hw):
                        n *= s
                    off += (n * x.element_size() + 255) & ~255
            slot_bytes = off
            if self.data_top + 2 * slot_bytes > self.arena_bytes or self.flag_top + self.FLAGS_PER_LAYER > self.nflags:
                raise _lib.SpconvError("halo mailbox arena exhausted; raise SPCONV_ARENA_MB")
            slots[key] = dict(data=self.data_top, flags=self.flag_top, offs=offs, slot_bytes=slot_bytes, seq=0)
            self.data_top += 2 * slot_bytes
            self.flag_top += self.FLAGS_PER_LAYER
        return slots[key]

    def exchange(self, layer, x, hh, hw, mask, ranks):
        """Two kernels per exchange: spc_halo_post (wait acks -> pack into the neighbours' slots ->
        signal) and spc_halo_collect (wait arrivals -> copy the strips out -> ack)."""
        L = _lib.lib()
        st = _stream()
        slot = self._slot(layer, x, hh, hw)
        slot["seq"] += 1
        seq = slot["seq"]
        par = seq & 1
        dirs = [i for i in range(9) if i != 4 and mask[i]]
        fb = slot["flags"]
        base_off = slot["data"] + par * slot["slot_bytes"]
        N, Cc, H, W = x.shape
        P9, I9, S9 = C.c_void_p * 9, C.c_int * 9, C.c_size_t * 9
        send, peers, src, dst, nbytes = P9(), P9(), P9(), P9(), S9()
        ack_local, arr_peer, arr_local, ack_peer = I9(), I9(), I9(), I9()
        recv = [None] * 9
        for d in dirs:
            pmb, pbase = self._peer(ranks[d])
            peers[d] = pmb
            send[d] = pbase + base_off + slot["offs"][8 - d]      # my strip d is the neighbour's strip 8-d
            ack_local[d] = fb + 18 + par * 9 + d                   # neighbour acks what I wrote (my flag)
            arr_peer[d] = fb + par * 9 + (8 - d)                   # I announce it on the neighbour's flag
            arr_local[d] = fb + par * 9 + d                        # neighbour announces my strip d here
            ack_peer[d] = fb + 18 + par * 9 + (8 - d)              # and I ack on its flag
            shp = strip_shape(d, N, Cc, H, W, hh, hw)
            recv[d] = torch.empty(shp, dtype=x.dtype, device=x.device)
            dst[d] = recv[d].data_ptr()
            src[d] = self.base + base_off + slot["offs"][d]
            nbytes[d] = recv[d].numel() * recv[d].element_size()
        _lib.check(L.spc_halo_post(N, Cc, H, W, hh, hw, _lib.dtype_code(x.dtype), C.c_void_p(x.data_ptr()), C.byref(send),
                                   self.mb, C.byref(peers), C.byref(ack_local), seq - 2 if seq > 2 else 0,
                                   C.byref(arr_peer), seq, st), "spc_halo_post")
        _lib.check(L.spc_halo_collect(C.byref(dst), C.byref(src), C.byref(nbytes), self.mb, C.byref(peers),
                                      C.byref(arr_local), seq, C.byref(ack_peer), st), "spc_halo_collect")
        return recv


_transport = None


def get_transport(device):
    global _transport
    if _transport is None:
        kind = os.environ.get("SPCONV_HALO_TRANSPORT", "peer" if device.type == "cuda" else "dist")
        _transport = PeerTransport(device) if kind == "peer" else DistTransport()
    return _transport


def overlap_enabled():
    """Overlap the exchange (comm stream) with the interior pass; SPCONV_HALO_OVERLAP=0 serialises."""
    return os.environ.get("SPCONV_HALO_OVERLAP", "1") != "0"


def set_transport(t):
    global _transport
    _transport = t
