"""D-slab mode: ONE cost volume split along depth over the GPUs of a box (BASELINE config 5).

Rank r owns planes [r*D/G, (r+1)*D/G) of every tensor of the path.  The kernels are the single-GPU ones, run on a
depth window (include/mvsnet_b200.h, mvsb200_slab_*); this module is the exchange step between them, over
`torch.distributed` (NCCL on NVLink / NVSwitch; gloo in the CPU tests):

  * cost volume: no exchange -- every rank holds all feature maps and computes its own two halo planes;
  * after every layer: all-reduce (SUM) of its batch statistics (BN uses statistics of the WHOLE volume,
    network.py:496) and a halo exchange of the boundary planes of its output with both neighbours (the 3x3x3
    receptive field of the consumers, network.py:210,327);
  * after the last layer: the softmax over depth (model.py:474) is split as well -- per-rank (max, sum of exp,
    depth-weighted sum) partials, one 3-map all-gather, a combine, and a sum of the per-rank shares of the
    probability map.

`exchange_layer` only touches byte regions of a flat workspace tensor, so it is testable on CPU.

With p2p=True the exchange moves INTO the kernels (no NCCL call between layers): the slab workspaces are CUDA-IPC
mapped into every rank, the producing epilogue stores its boundary planes straight into the neighbours' halo planes
over NVLink, a one-block kernel publishes the layer's statistics and raises a flag on every rank, and the consuming
kernel waits on its local flags (mvsb200_slab_layer_p2p).  Only the small collectives of the regression remain; they
are also the barrier that keeps a fast rank's next inference out of a slow rank's buffers.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

N_LAYERS = 11


def slab_range(depth_num: int, rank: int, world: int):
    """[begin, end) of the depth planes rank `rank` owns; the regularizer needs multiples of 8 planes per slab."""
    if depth_num % world or (depth_num // world) % 8:
        raise ValueError(f"depth {depth_num} does not split into {world} slabs of a multiple of 8 planes")
    dl = depth_num // world
    return rank * dl, (rank + 1) * dl


def layer_regions(layer: int, n_views: int, depth_num: int, world: int, hf: int, wf: int, channels: int,
                  base_filter: int) -> dict:
    """Byte regions of the slab workspace the host exchanges after `layer` (host-only call)."""
    lib = L.load()
    out = (ctypes.c_ulonglong * 14)()
    L.check(lib.mvsb200_slab_regions(layer, n_views, depth_num, world, hf, wf, channels, base_filter, out), "slab_regions")
    v = [int(x) for x in out]
    tensors = []
    for t in range(2):
        plane, first, last, before, after = v[2 + 5 * t:7 + 5 * t]
        if plane:
            tensors.append(dict(plane=plane, first=first, last=last, before=before, after=after))
    return dict(stats=(v[0], v[1]), tensors=tensors, filtered=(v[12], v[13]))


# RegNetUS0 data flow (mvsnetworks.py:131-158): producers whose exchanged outputs / statistics layer i reads
LAYER_INPUTS = {0: [], 1: [0], 2: [1], 3: [], 4: [0], 5: [1], 6: [2], 7: [6], 8: [7, 5], 9: [8, 4], 10: [9, 3]}
# Execution order in slab mode: every layer directly after a layer it does NOT depend on where possible, so that
# the exchange of one layer's output runs under the next layer's kernels (3dconv1_0 | 3dconv0_1 | 3dconv2_0 | ...)
SLAB_ORDER = [0, 3, 1, 4, 2, 5, 6, 7, 8, 9, 10]


def exchange_layer(ws: torch.Tensor, regions: dict, rank: int, world: int, group=None, wait: bool = True):
    """All-reduce the layer's statistics and swap boundary planes with the neighbours.  `ws` is the flat uint8
    workspace; everything happens in place.  With wait=False the pending work handles are returned and the caller
    waits on them before the first consumer of the layer runs."""
    pending = []
    off, nbytes = regions["stats"]
    if nbytes:
        w = dist.all_reduce(ws[off:off + nbytes].view(torch.float64), op=dist.ReduceOp.SUM, group=group, async_op=True)
        pending.append(w)
    opsl: List[dist.P2POp] = []
    for t in regions["tensors"]:
        n = t["plane"]
        if rank > 0:          # my first plane is the previous rank's AFTER halo; its last plane is my BEFORE halo
            opsl.append(dist.P2POp(dist.isend, ws[t["first"]:t["first"] + n], rank - 1, group))
            opsl.append(dist.P2POp(dist.irecv, ws[t["before"]:t["before"] + n], rank - 1, group))
        if rank < world - 1:
            opsl.append(dist.P2POp(dist.isend, ws[t["last"]:t["last"] + n], rank + 1, group))
            opsl.append(dist.P2POp(dist.irecv, ws[t["after"]:t["after"] + n], rank + 1, group))
    if opsl:
        pending.extend(dist.batch_isend_irecv(opsl))
    if wait:
        for w in pending:
            w.wait()
        return []
    return pending


class DSlabHotPath:
    """feats [N,Hf,Wf,32] + cams -> depth map + probability map with the volume's depth split over the ranks of
    `group`.  Every rank calls infer() with the same inputs and gets the same maps."""

    def __init__(self, n_views, depth_num, hf, wf, weights, channels=32, order="mem", inverse_depth=False, bn_eps=1e-5,
                 device="cuda", group=None, p2p=False):
        from .engine import RegnetWeights
        self.lib = L.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.begin, self.end = slab_range(depth_num, self.rank, self.world)
        self.device = torch.device(device)
        self.n_views, self.depth_num, self.hf, self.wf, self.channels = n_views, depth_num, hf, wf, channels
        self.order = ops._ORDER[order]
        self.inverse_depth = int(bool(inverse_depth))
        self.bn_eps = float(bn_eps)
        self.weights = weights if isinstance(weights, RegnetWeights) else RegnetWeights(weights, self.device)
        self.base_filter = self.weights.base_filter
        nbytes = self.lib.mvsb200_slab_workspace_bytes(n_views, depth_num, self.world, hf, wf, channels, self.base_filter)
        if nbytes == 0:
            raise L.MVSB200Error("slab_workspace_bytes rejected the shape: " + L.last_error())
        self.p2p = bool(p2p) and self.world > 1
        self.seq = 0
        if self.p2p:
            self._map_peers(nbytes)
        else:
            self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        self.regions = [layer_regions(i, n_views, depth_num, self.world, hf, wf, channels, self.base_filter)
                        for i in range(N_LAYERS)]
        self.partial = torch.empty((3, hf * wf), dtype=torch.float32, device=self.device)
        self.partials = torch.empty((self.world, 3, hf * wf), dtype=torch.float32, device=self.device)

    def _map_peers(self, nbytes: int):
        """Slab workspace in plain cudaMalloc memory, exported over CUDA IPC and opened by every other rank."""
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            L.check(self.lib.mvsb200_ipc_alloc(nbytes, ctypes.byref(base)), "ipc_alloc")
            handle = ctypes.create_string_buffer(64)
            L.check(self.lib.mvsb200_ipc_export(base, handle), "ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
            addrs = []
            for q in range(self.world):
                if q == self.rank:
                    addrs.append(base.value)
                else:
                    peer = ctypes.c_void_p()
                    L.check(self.lib.mvsb200_ipc_open(ctypes.create_string_buffer(handles[q], 64), ctypes.byref(peer)),
                            "ipc_open")
                    addrs.append(peer.value)
        self._base, self._addrs, self._nbytes = base, addrs, nbytes
        self.peers_host = (ctypes.c_void_p * self.world)(*addrs)
        self.peers_dev = torch.tensor(addrs, dtype=torch.int64, device=self.device)

        class _Raw:                                   # zero-copy torch view of the library-owned allocation
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (base.value, False), "version": 2}
        self.ws = torch.as_tensor(_Raw(), device=self.device)

    def close(self):
        """Unmap the peers and free the IPC workspace (p2p mode); call it on every rank after a barrier."""
        if getattr(self, "_base", None) is not None:
            torch.cuda.synchronize(self.device)
            for q, a in enumerate(self._addrs):
                if q != self.rank:
                    self.lib.mvsb200_ipc_close(ctypes.c_void_p(a))
            self.ws = None
            self.lib.mvsb200_ipc_free(self._base)
            self._base = None

    def abort(self):
        """p2p mode: release this rank's kernels that wait for another rank's publication flag (a watchdog that has
        lost a rank calls this on the survivors; the next p2p_error() reports it).  Safe from another host thread."""
        if self.p2p:
            with torch.cuda.device(self.device):
                L.check(self.lib.mvsb200_slab_p2p_abort(self.n_views, self.depth_num, self.world, self.hf, self.wf,
                                                        self.channels, self.base_filter, L.ptr(self.ws)), "slab_p2p_abort")

    def p2p_error(self) -> bool:
        """True when a kernel of the last inference gave up waiting (time-out or abort()); clears the flags."""
        if not self.p2p:
            return False
        with torch.cuda.device(self.device):
            return bool(self.lib.mvsb200_slab_p2p_error(self.n_views, self.depth_num, self.world, self.hf, self.wf,
                                                        self.channels, self.base_filter, L.ptr(self.ws), L.stream_ptr()))

    def infer(self, feats: torch.Tensor, cams: torch.Tensor, depth_start: float, depth_interval: float):
        L.require_cuda(feats, cams)
        args = (self.n_views, self.depth_num, self.rank, self.world, self.hf, self.wf, self.channels)
        rc = self.lib.mvsb200_slab_begin(L.ptr(feats.contiguous()), L.ptr(cams.contiguous()), *args,
                                         float(depth_start), float(depth_interval), self.inverse_depth, self.order,
                                         ctypes.byref(self.weights.params), self.base_filter, L.ptr(self.ws),
                                         self.ws.numel(), L.stream_ptr())
        L.check(rc, "slab_begin")
        if self.p2p:
            self.seq += 1
            for layer in SLAB_ORDER:
                rc = self.lib.mvsb200_slab_layer_p2p(layer, *args, ctypes.byref(self.weights.params), self.base_filter,
                                                     self.bn_eps, L.ptr(self.ws), L.ptr(self.peers_dev), self.peers_host,
                                                     self.seq, L.stream_ptr())
                L.check(rc, f"slab_layer_p2p {layer}")
            return self._finish(depth_start, depth_interval)
        pending = {}
        for layer in SLAB_ORDER:
            for src in LAYER_INPUTS[layer]:                 # the exchanges this layer reads must have landed
                for w in pending.pop(src, []):
                    w.wait()
            rc = self.lib.mvsb200_slab_layer(layer, *args, ctypes.byref(self.weights.params), self.base_filter,
                                             self.bn_eps, L.ptr(self.ws), L.stream_ptr())
            L.check(rc, f"slab_layer {layer}")
            if layer != N_LAYERS - 1:
                pending[layer] = exchange_layer(self.ws, self.regions[layer], self.rank, self.world, self.group,
                                                wait=False)
        for works in pending.values():
            for w in works:
                w.wait()
        return self._finish(depth_start, depth_interval)

    def _finish(self, depth_start: float, depth_interval: float):
        """Softmax over depth across the slabs: per-rank partials, one small all-gather, combine, and a sum of the
        per-rank shares of the probability map (no rank ever holds the whole filtered volume)."""
        off, nbytes = self.regions[N_LAYERS - 1]["filtered"]
        mine = self.ws[off:off + nbytes].view(torch.float32)
        npix, dl = self.hf * self.wf, self.end - self.begin
        rc = self.lib.mvsb200_regress_partial(L.ptr(mine), dl, self.begin, self.depth_num, npix, float(depth_start),
                                              float(depth_interval), self.inverse_depth, L.ptr(self.partial),
                                              L.stream_ptr())
        L.check(rc, "regress_partial")
        dist.all_gather_into_tensor(self.partials.view(-1), self.partial.view(-1), group=self.group)
        depth = torch.empty((self.hf, self.wf), dtype=torch.float32, device=self.device)
        prob = torch.empty((self.hf, self.wf), dtype=torch.float32, device=self.device)
        rc = self.lib.mvsb200_regress_combine(L.ptr(self.partials), self.world, L.ptr(mine), dl, self.begin,
                                              self.depth_num, npix, float(depth_start), float(depth_interval),
                                              self.inverse_depth, 4, L.ptr(depth), L.ptr(prob), L.stream_ptr())
        L.check(rc, "regress_combine")
        dist.all_reduce(prob, op=dist.ReduceOp.SUM, group=self.group)
        return depth, prob


class LocalSlabHotPath:
    """The D-slab path with every slab on ONE GPU, in ONE process: the same library calls as DSlabHotPath
    (mvsb200_slab_begin / _layer / _regress_partial / _combine on `slabs` workspaces), with the exchange step done by
    device copies between the workspaces and the statistics summed in place.  No torch.distributed: this is how a
    single-GPU box checks the slab kernels (depth windows, halo planes, global batch statistics, split softmax) against
    the oracle; it is not a fast path."""

    def __init__(self, n_views, depth_num, hf, wf, weights, slabs, channels=32, order="mem", inverse_depth=False,
                 bn_eps=1e-5, device="cuda"):
        from .engine import RegnetWeights
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.world = int(slabs)
        slab_range(depth_num, 0, self.world)                      # validates the split
        self.n_views, self.depth_num, self.hf, self.wf, self.channels = n_views, depth_num, hf, wf, channels
        self.order = ops._ORDER[order]
        self.inverse_depth = int(bool(inverse_depth))
        self.bn_eps = float(bn_eps)
        self.weights = weights if isinstance(weights, RegnetWeights) else RegnetWeights(weights, self.device)
        self.base_filter = self.weights.base_filter
        nbytes = self.lib.mvsb200_slab_workspace_bytes(n_views, depth_num, self.world, hf, wf, channels, self.base_filter)
        if nbytes == 0:
            raise L.MVSB200Error("slab_workspace_bytes rejected the shape: " + L.last_error())
        self.ws = [torch.zeros((nbytes,), dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        self.regions = [layer_regions(i, n_views, depth_num, self.world, hf, wf, channels, self.base_filter)
                        for i in range(N_LAYERS)]

    def _exchange(self, layer: int):
        reg = self.regions[layer]
        off, nbytes = reg["stats"]
        if nbytes:
            total = sum(w[off:off + nbytes].view(torch.float64) for w in self.ws)
            for w in self.ws:
                w[off:off + nbytes].view(torch.float64).copy_(total)
        for t in reg["tensors"]:
            n = t["plane"]
            for r in range(self.world):
                if r > 0:        # my first plane is the previous slab's AFTER halo
                    self.ws[r - 1][t["after"]:t["after"] + n].copy_(self.ws[r][t["first"]:t["first"] + n])
                if r < self.world - 1:
                    self.ws[r + 1][t["before"]:t["before"] + n].copy_(self.ws[r][t["last"]:t["last"] + n])

    def infer(self, feats: torch.Tensor, cams: torch.Tensor, depth_start: float, depth_interval: float):
        with torch.cuda.device(self.device):
            L.require_cuda(feats, cams)
            feats, cams = feats.contiguous(), cams.contiguous()
            common = (self.n_views, self.depth_num)
            shape = (self.hf, self.wf, self.channels)
            for r in range(self.world):
                rc = self.lib.mvsb200_slab_begin(L.ptr(feats), L.ptr(cams), *common, r, self.world, *shape,
                                                 float(depth_start), float(depth_interval), self.inverse_depth, self.order,
                                                 ctypes.byref(self.weights.params), self.base_filter, L.ptr(self.ws[r]),
                                                 self.ws[r].numel(), L.stream_ptr())
                L.check(rc, "slab_begin")
            for layer in SLAB_ORDER:
                for r in range(self.world):
                    rc = self.lib.mvsb200_slab_layer(layer, *common, r, self.world, *shape,
                                                     ctypes.byref(self.weights.params), self.base_filter, self.bn_eps,
                                                     L.ptr(self.ws[r]), L.stream_ptr())
                    L.check(rc, f"slab_layer {layer}")
                if layer != N_LAYERS - 1:
                    self._exchange(layer)
            # softmax over depth across the slabs (model.py:474): per-slab partials, combine, summed probability shares
            npix, dl = self.hf * self.wf, self.depth_num // self.world
            off, nbytes = self.regions[N_LAYERS - 1]["filtered"]
            partials = torch.empty((self.world, 3, npix), dtype=torch.float32, device=self.device)
            for r in range(self.world):
                mine = self.ws[r][off:off + nbytes].view(torch.float32)
                rc = self.lib.mvsb200_regress_partial(L.ptr(mine), dl, r * dl, self.depth_num, npix, float(depth_start),
                                                      float(depth_interval), self.inverse_depth, L.ptr(partials[r]),
                                                      L.stream_ptr())
                L.check(rc, "regress_partial")
            depth = torch.empty((self.hf, self.wf), dtype=torch.float32, device=self.device)
            prob = torch.zeros((self.hf, self.wf), dtype=torch.float32, device=self.device)
            share = torch.empty_like(prob)
            for r in range(self.world):
                mine = self.ws[r][off:off + nbytes].view(torch.float32)
                rc = self.lib.mvsb200_regress_combine(L.ptr(partials), self.world, L.ptr(mine), dl, r * dl, self.depth_num,
                                                      npix, float(depth_start), float(depth_interval), self.inverse_depth,
                                                      4, L.ptr(depth), L.ptr(share), L.stream_ptr())
                L.check(rc, "regress_combine")
                prob += share
            return depth, prob

    def filtered_volume(self) -> torch.Tensor:
        """The filtered cost volume [D,Hf,Wf] fp32 of the last infer(), assembled from the slabs."""
        off, nbytes = self.regions[N_LAYERS - 1]["filtered"]
        dl = self.depth_num // self.world
        return torch.cat([w[off:off + nbytes].view(torch.float32).view(dl, self.hf, self.wf) for w in self.ws], dim=0)
