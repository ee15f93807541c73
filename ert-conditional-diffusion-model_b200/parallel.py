"""Multi-GPU ensemble: members are independent (ECD.py:106-119 has no cross-row operation), so
each rank runs a contiguous slice of the ensemble and ONE all-gather returns the fields in
member order (SURVEY.md §8 e).  One process per GPU, ``torch.distributed`` (NCCL) as plumbing.

The slicing keeps results independent of the number of ranks: device-RNG streams are keyed by
the GLOBAL member index (``member_offset``) and injected noise is read from the member-slice of
the full tensor, so 1, 2, 4 or 8 ranks give bit-identical gathered fields.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _RawCuda:
    """A device allocation of the library seen as a CUDA array (zero-copy view for torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class PeerAllGather:
    """All-gather over NVLink peer memory through the library's own kernel (``ertdiff_peer_*``) instead of NCCL:
    one launch stores this rank's slice into every peer's buffer (P2P stores), publishes an epoch flag and waits for
    the peers' flags.  For the path's two latency-bound collectives (the final fields, the packed statistics
    records) on the GPUs of one node, one process per GPU.  ``torch.distributed`` is used once, at set-up, to
    exchange the CUDA IPC handles.

    ``all_gather(src)`` returns a view of this rank's gathered buffer ``(world * slot,) + src.shape[1:]`` that stays
    valid until the next-but-one call; every rank must call in the same order with equal shapes."""

    def __init__(self, max_bytes_per_rank: int, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.slot_cap = (int(max_bytes_per_rank) + 255) // 256 * 256
        lib = _lib.load()
        handle = (C.c_ubyte * 64)()
        self._h = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()

        def agree(err, what):
            # set-up is collective: if a step fails on ANY rank, every rank raises (nobody is left waiting in the next one)
            ok = torch.tensor([0.0 if err else 1.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0:
                self.close()
                raise _lib.ErtdiffError(f"PeerAllGather: {what} failed on at least one rank"
                                        + (f" (this rank: {err})" if err else ""))

        err = None
        try:
            _lib.check(lib.ertdiff_peer_create(C.byref(self._h), index, self.rank, self.world,
                                               self.slot_cap * self.world, handle), "peer_create")
        except Exception as exc:
            err = exc
        agree(err, "allocating / exporting the peer buffer (CUDA IPC)")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        blob = bytes(every.cpu().numpy().tobytes())
        try:
            _lib.check(lib.ertdiff_peer_connect(self._h, blob), "peer_connect")
        except Exception as exc:
            err = exc
        agree(err, "mapping the peers' buffers (cudaIpcOpenMemHandle)")      # also: every buffer is mapped before the first store

    def all_gather(self, src: torch.Tensor) -> torch.Tensor:
        C, lib = self._C, self._lib.load()
        src = src.contiguous()
        nbytes = src.numel() * src.element_size()
        if nbytes > self.slot_cap:
            raise ValueError(f"slice of {nbytes} bytes exceeds the peer buffer's {self.slot_cap} per rank")
        slot = (nbytes + 15) // 16 * 16
        out = C.c_void_p()
        with torch.cuda.device(self.device):
            self._lib.check(lib.ertdiff_peer_all_gather(self._h, self._lib.ptr(src), nbytes, slot, C.byref(out),
                                                        self._lib.stream_ptr(self.device)), "peer_all_gather")
        raw = torch.as_tensor(_RawCuda(out.value, slot * self.world), device=self.device)
        if slot == nbytes:
            return raw.view(src.dtype).view((self.world * src.size(0),) + tuple(src.shape[1:]))
        per = raw.view(self.world, slot)[:, :nbytes].contiguous()
        return per.view(src.dtype).view((self.world * src.size(0),) + tuple(src.shape[1:]))

    def status(self) -> int:
        st = self._C.c_int()
        self._lib.check(self._lib.load().ertdiff_peer_status(self._h, self._C.byref(st)), "peer_status")
        return int(st.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.load().ertdiff_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def member_slice(n_members: int, rank: int, world_size: int, multiple_of: int = 1):
    """Contiguous ``[start, stop)`` of rank's members; slice sizes are multiples of
    ``multiple_of`` (the number of distinct conditions, so ``member % n_cond`` is preserved)."""
    if n_members % multiple_of:
        raise ValueError("n_members must be a multiple of multiple_of")
    groups = n_members // multiple_of
    base, extra = divmod(groups, world_size)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start * multiple_of, stop * multiple_of


def gather_members(x_local: torch.Tensor, n_members: int, multiple_of: int = 1,
                   group=None, peer: "PeerAllGather | None" = None) -> torch.Tensor:
    """All-gather ``(B_local, P)`` slices into ``(n_members, P)`` in rank (= member) order.
    Slices may differ in length by one group: each rank pads to the longest slice, one
    equal-sized all-gather runs, and the padding is dropped."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_local
    world = dist.get_world_size(group)
    counts = [b - a for a, b in (member_slice(n_members, r, world, multiple_of)
                                 for r in range(world))]
    longest = max(counts)
    x_local = x_local.contiguous()
    even = min(counts) == longest
    if not even and x_local.size(0) < longest:
        pad = torch.zeros((longest - x_local.size(0),) + tuple(x_local.shape[1:]),
                          device=x_local.device, dtype=x_local.dtype)
        x_local = torch.cat([x_local, pad], dim=0)
    if even and x_local.is_cuda and peer is not None:
        return peer.all_gather(x_local)                                # the library's own NVLink kernel
    if even and x_local.is_cuda:
        out = torch.empty((n_members,) + tuple(x_local.shape[1:]), device=x_local.device,
                          dtype=x_local.dtype)
        dist.all_gather_into_tensor(out, x_local, group=group)     # one NCCL all-gather
        return out
    parts = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(parts, x_local, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def sample_ensemble_sharded(chain_fn, n_members: int, n_cond: int = 1, group=None, noise=None):
    """Run ``chain_fn(start, stop, noise_slice) -> (stop-start, P)`` on this rank's slice and
    gather.  ``chain_fn`` is normally a closure over ``sampler.run_chain`` with
    ``member_offset=start``; tests inject a CPU stand-in to exercise the slicing and the
    collective on the gloo backend.  With ``n_cond`` distinct conditions slices are whole
    realisations (multiples of ``n_cond``), so ``member % n_cond`` is the same on every rank."""
    start, stop = member_slice(n_members, dist.get_rank(group) if dist.is_initialized() else 0,
                               dist.get_world_size(group) if dist.is_initialized() else 1, n_cond)
    nz = None if noise is None else noise[:, start:stop, :]
    x_local = chain_fn(start, stop, nz)
    return gather_members(x_local, n_members, n_cond, group)


def column_slice(n_columns: int, rank: int, world_size: int):
    """Contiguous ``[start, stop)`` of the columns (pixels / parameters) whose statistics this rank
    computes."""
    return member_slice(n_columns, rank, world_size, 1)


_UNPAD_INDEX = {}      # (Q, world, device) -> row indices that drop the padding of the gathered (world*q_max) records


def ensemble_statistics_distributed(x: torch.Tensor, percentiles=(25, 50, 75), n_grid: int = 5000, group=None, shard=None,
                                    stats_fn=None, peer: "PeerAllGather | None" = None):
    """Statistics of the gathered fields ``x (N, Q)`` on every rank: with more than one rank the columns are split over
    the ranks (``sharded_statistics``), otherwise the one-call summary runs on the whole array.  ``shard=False`` makes
    every rank compute all columns itself (no second collective) -- measured SLOWER on 8 GPUs even for 2048
    gathered members (0.518 vs 0.456 ms per step): the KDE scan of 29 columns is throughput-bound (115 us against 30
    for a rank's 4 columns), which outweighs the ~20 us of the packed all-gather.  Same bits either way."""
    from . import stats as st
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if shard is None:
        shard = world > 1
    if shard:
        return sharded_statistics(x, percentiles, n_grid, group, stats_fn, peer)
    if stats_fn is not None:            # (CPU tests)
        return st.summary_views(stats_fn(x, None).t().contiguous(), len(percentiles))
    return st.ensemble_summary(x, percentiles, n_grid)


def sharded_statistics(x: torch.Tensor, percentiles=(25, 50, 75), n_grid: int = 5000, group=None,
                       stats_fn=None, peer: "PeerAllGather | None" = None):
    """Ensemble statistics of the gathered fields ``x (N, Q)`` with the COLUMNS split over the ranks:
    every statistic of the path is per column (ECD.py:747-762, 867-872), so rank r computes columns
    ``column_slice(Q, r, world)`` over all N members and one small all-gather returns every map to
    every rank.  The KDE grid spans the global min/max of the whole array (ECD.py:749-751), which each
    rank takes from its own copy of ``x`` -- so the result is bit-identical to the unsharded call for
    any number of ranks.  Returns ``{"mean","std","var","pct" (len(percentiles), Q),"mode","mode_index"}``
    as float64 views of one gathered block ``"packed" (Q, 5 + len(percentiles))`` (``mode_index``: integral values).

    Per rank and step: ONE library call (``ertdiff_ensemble_summary``: the statistics kernels on its column window,
    packed as one float64 record per column), ONE equal-sized all-gather of ``(q_max, rows)`` records, and (only
    when ``Q`` does not divide evenly) one gather that drops the padding records.
    ``stats_fn(x_cols, lohi) -> (rows, q_local)`` float64 replaces the device kernels in CPU tests."""
    from . import stats as st
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N, Q = x.shape
    a, b = column_slice(Q, rank, world)
    nq = len(percentiles)
    rows = 5 + nq
    q_max = -(-Q // world)
    block = torch.zeros(q_max, rows, device=x.device, dtype=torch.float64) if (b - a) < q_max else \
        torch.empty(q_max, rows, device=x.device, dtype=torch.float64)
    if b > a:
        if stats_fn is None:
            st.ensemble_summary_packed(x, percentiles, n_grid, col0=a, ncols=b - a, out=block)
        else:
            block[:b - a].copy_(stats_fn(x[:, a:b].contiguous(), None).t())
    if world > 1:
        gathered = None
        if x.is_cuda and peer is not None:
            gathered = peer.all_gather(block)                              # the library's own NVLink kernel
        elif x.is_cuda:
            gathered = torch.empty(world * q_max, rows, device=x.device, dtype=torch.float64)
            dist.all_gather_into_tensor(gathered, block, group=group)      # one NCCL all-gather
        else:
            parts = [torch.empty_like(block) for _ in range(world)]
            dist.all_gather(parts, block, group=group)
            gathered = torch.cat(parts, dim=0)
        if Q % world:
            key = (Q, world, str(x.device))
            index = _UNPAD_INDEX.get(key)
            if index is None:
                keep = []
                for r in range(world):
                    ra, rb = column_slice(Q, r, world)
                    keep.extend(range(r * q_max, r * q_max + (rb - ra)))
                index = torch.tensor(keep, device=x.device, dtype=torch.int64)
                _UNPAD_INDEX[key] = index
            gathered = gathered.index_select(0, index)
        full = gathered
    else:
        full = block[:Q]
    return st.summary_views(full, nq)


def sharded_misfit(sim_local: torch.Tensor, observed: torch.Tensor, n_maps: int, A: float = 0.1, B: float = 0.01,
                   group=None, misfit_fn=None):
    """Per-map misfit metrics (ECD.py:764-786, 927-930) with the simulated MAPS split over the ranks: each
    map's WSSE / MSE depends on that map alone, so rank r evaluates its ``member_slice(n_maps, r, world)``
    maps ``sim_local (n_local, L, C)`` against the replicated ``observed (L, C)`` and one all-gather returns
    every value to every rank in map order; the ranking (argsort of the totals) is then taken from the
    gathered totals, so every result is identical to the unsharded ``stats.misfit_metrics`` for any number of
    ranks.  ``misfit_fn(sim_local, observed) -> {"wsse","wsse_total","mse"}`` replaces the device kernels in
    CPU tests."""
    from . import stats as st
    if misfit_fn is None:
        def misfit_fn(s, o):
            return st.misfit_metrics(s, o, A=A, B=B)
    C = observed.shape[-1]
    if sim_local.size(0):
        r = misfit_fn(sim_local, observed)
        packed = torch.cat([r["wsse"], r["wsse_total"][:, None], r["mse"][:, None]], dim=1)     # (n_local, C + 2)
    else:
        packed = torch.zeros(0, C + 2, device=sim_local.device, dtype=sim_local.dtype)
    full = gather_members(packed, n_maps, 1, group)
    total = full[:, C].contiguous()
    # (CPU tensors only reach this point from the gloo tests' stand-in misfit_fn)
    order = st.argsort_stable(total) if total.is_cuda else torch.argsort(total, stable=True)
    return {"wsse": full[:, :C].contiguous(), "wsse_total": total, "order": order,
            "mse": full[:, C + 1].contiguous()}
