"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  SURVEY.md 8(e):

  * a large sheet is split into contiguous bands over the slow axis n2 ("row bands"); the only
    data-path exchange is the R clean-window halo lines of packed RGB between band neighbours
    (grouped isend/irecv = ncclSend/ncclRecv), after which every rank runs ONE fused launch;
  * a batch of sheets (config 5) is split into whole sheets per rank, no communication.

The reference has no distributed code at all (single-threaded Julia); this module only exists
for the new path.  Band buffers have shape (R + nb + R, n1, bpp): own lines in the middle, halo
slots at both ends, so received halos land exactly where the kernel reads them.
"""
import torch
import torch.distributed as dist


def band_range(n2, world, rank):
    """Lines [a, b) owned by `rank` (balanced contiguous split of the slow axis)."""
    return (rank * n2) // world, ((rank + 1) * n2) // world


def sheet_range(n_sheets, world, rank):
    return (rank * n_sheets) // world, ((rank + 1) * n_sheets) // world


def set_known_for_band(ctx, lin_idx, cat, n1, n2, R, rank, world):
    """Known-category points of the WHOLE sheet (lin_idx = i + j * n1, `_set_categories!`,
    src/selectcolors.jl:670-681) -> the points rank's band can see (its own lines and the R halo lines
    on both sides), uploaded with the band's first sheet line (mr_set_known_band).  Every rank calls this
    with the same lists; the following categorize_clean_band / PeerHalos.categorize_clean then applies the
    override of src/categories.jl:110-112 exactly as a whole-image call would."""
    import numpy as np
    a, b = band_range(n2, world, rank)
    idx = np.asarray(lin_idx, dtype=np.int64).reshape(-1)
    ct = np.asarray(cat, dtype=np.uint8).reshape(-1)
    j = idx // max(n1, 1)
    sel = (j >= a - R) & (j < b + R)
    ctx.set_known_band(idx[sel], ct[sel], a)
    return int(sel.sum())


def alloc_band(n1, nb, R, bpp=3, device="cuda"):
    """Band buffer with halo slots; `own` is the view of the rank's own lines."""
    buf = torch.empty((nb + 2 * R, n1, bpp), dtype=torch.uint8, device=device)
    return buf, buf[R:R + nb]


def exchange_halos(buf, R, rank=None, world=None, group=None):
    """Fill the halo slots of `buf` with the neighbours' edge lines.  Returns (halo_lo, halo_hi):
    the number of real lines now readable before / after the own lines (0 at the sheet border)."""
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    nb = buf.shape[0] - 2 * R
    if R == 0 or world == 1:
        return 0, 0
    if nb < R:
        raise ValueError(f"band of {nb} lines is thinner than the window radius {R}")
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, buf[R:2 * R].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[0:R], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, buf[nb:nb + R].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[nb + R:nb + 2 * R], rank + 1, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return (R if rank > 0 else 0), (R if rank < world - 1 else 0)


def categorize_clean_band(ctx, buf, R, out, rank=None, world=None, group=None, stream=0):
    """Halo exchange + one fused launch on this rank's band.  buf: (R+nb+R, n1, bpp) uint8 CUDA
    tensor, out: (nb, n1) uint8 CUDA tensor.  Asynchronous on `stream` (after the exchange)."""
    nb = buf.shape[0] - 2 * R
    n1, bpp = buf.shape[1], buf.shape[2]
    halo_lo, halo_hi = exchange_halos(buf, R, rank, world, group)
    own = buf[R:R + nb]
    ctx.categorize_clean_dev(own, n1, nb, n1 * bpp, bpp, R, out, n1, halo_lo, halo_hi, stream)
    return halo_lo, halo_hi


class PeerHalos:
    """Band neighbours' edge lines mapped into this process (CUDA IPC over NVLink peer memory), so
    that the fused kernel reads the 2R halo lines in place from the neighbours' HBM and a band-
    sharded sheet needs no exchange step at all.  One instance per band buffer; collective over
    `group` (handles are all-gathered once).  The band buffers must stay allocated while in use."""

    def __init__(self, ctx, buf, R, rank=None, world=None, group=None):
        from . import _lib
        self.ctx, self.R = ctx, R
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        nb = buf.shape[0] - 2 * R
        stride = buf.stride(0) * buf.element_size()
        handle, off = _lib.ipc_export(buf)
        mine = (handle, off, nb, stride)
        infos = [None] * self.world
        dist.all_gather_object(infos, mine, group=group)
        self._bases = []
        self.lo_ptr = self.hi_ptr = 0
        self.halo_lo = self.halo_hi = 0
        if R > 0 and self.rank > 0:                       # lower neighbour's last R own lines
            h, o, nbp, st = infos[self.rank - 1]
            base = ctx.ipc_open(h)
            self._bases.append(base)
            self.lo_ptr, self.halo_lo = base + o + nbp * st, R      # own lines start at slot R: last R = slots nb .. nb+R-1
        if R > 0 and self.rank < self.world - 1:          # upper neighbour's first R own lines
            h, o, nbp, st = infos[self.rank + 1]
            base = ctx.ipc_open(h)
            self._bases.append(base)
            self.hi_ptr, self.halo_hi = base + o + R * st, R
        dist.barrier(group=group)

    def categorize_clean(self, buf, out, stream=0):
        """ONE fused launch on this rank's band; no exchange.  The neighbours' lines must be in place."""
        R = self.R
        nb = buf.shape[0] - 2 * R
        n1, bpp = buf.shape[1], buf.shape[2]
        self.ctx.categorize_clean_dev_peer(buf[R:R + nb], n1, nb, n1 * bpp, bpp, R, self.lo_ptr or None, self.halo_lo,
                                           self.hi_ptr or None, self.halo_hi, out, n1, stream)

    def close(self):
        for b in self._bases:
            self.ctx.ipc_close(b)
        self._bases = []
