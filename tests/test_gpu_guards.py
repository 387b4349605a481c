"""GPU (-m gpu): out-of-bounds WRITE detection without compute-sanitizer (closed on this pool: see
profiles/r02_sanitizer_closed.txt).  Every device entry point runs on buffers that sit inside larger
allocations filled with a canary pattern -- before the first line, after the last line and in the pitch
padding of every line; afterwards every canary byte must be intact, and the result must still equal the
oracle.  Geometries are chosen to hit the edge strips of the fast kernel (n1 not a multiple of the strip
width, of 32 and of 16), its TMA bulk stores (whole 16-byte chunks) and its byte patches, the repacking
path, the generic kernel, the segment prune, fill_gaps, blur and the peer-halo entry."""
import numpy as np
import pytest
import torch

from util import mismatches

pytestmark = pytest.mark.gpu
CANARY = 0xA5
GUARD = 4096        # bytes before and after


class Guarded:
    """n2 lines of `pitch` bytes, `used` bytes of each written by the callee; guards all around."""

    def __init__(self, n2, pitch, used, dev, fill=None):
        self.n2, self.pitch, self.used = n2, pitch, used
        self.buf = torch.full((GUARD + n2 * pitch + GUARD,), CANARY, dtype=torch.uint8, device=dev)
        self.view = self.buf[GUARD:GUARD + n2 * pitch].view(n2, pitch)
        if fill is not None:
            self.view[:, :used] = fill

    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def data(self):
        return self.view[:, :self.used].cpu().numpy()

    def check(self, what):
        assert bool((self.buf[:GUARD] == CANARY).all()), f"{what}: write before the buffer"
        assert bool((self.buf[GUARD + self.n2 * self.pitch:] == CANARY).all()), f"{what}: write after the buffer"
        if self.pitch > self.used:
            assert bool((self.view[:, self.used:] == CANARY).all()), f"{what}: write into the pitch padding"


GEOMS = [(1000, 61), (896, 40), (897, 33), (4113, 24), (16, 19), (31, 50), (2048, 17)]


@pytest.mark.parametrize("n1,n2", GEOMS)
@pytest.mark.parametrize("R", [0, 1, 2, 3])
def test_fused_and_clean_never_write_outside(ctx, mr, oracle, n1, n2, R):
    from maprast_b200 import synth
    dev = torch.device("cuda", 0)
    K = 16 if R != 3 else 40
    pal = synth.make_palette(300 + K, K)
    st = synth.make_stats(300 + K, pal)
    ctx.set_palette(st.mean, st.lo, st.hi)
    ctx.palette = None
    ctx.set_known(None)
    scan = oracle.synth_scan(21, 0, pal, n1, n2)
    want = oracle.categorize_clean(scan, st.mean, st.lo, st.hi, R)
    for in_pitch, out_pitch in ((n1 * 3, n1), ((n1 * 3 + 15) // 16 * 16 + 16, (n1 + 15) // 16 * 16 + 16)):
        src = Guarded(n2, in_pitch, n1 * 3, dev, torch.from_numpy(scan.reshape(n2, n1 * 3)).to(dev))
        dst = Guarded(n2, out_pitch, n1, dev)
        ctx.categorize_clean_dev(src.ptr(), n1, n2, in_pitch, 3, R, dst.ptr(), out_pitch)
        torch.cuda.synchronize()
        dst.check(f"fused R={R} pitches {in_pitch}/{out_pitch}")
        src.check("fused (input)")
        assert mismatches(dst.data(), want) == 0
        if R:
            lab = oracle.categorize_u8(scan, st.mean, st.lo, st.hi)
            lsrc = Guarded(n2, out_pitch, n1, dev, torch.from_numpy(lab).to(dev))
            ldst = Guarded(n2, out_pitch, n1, dev)
            ctx.clean_dev(lsrc.ptr(), n1, n2, out_pitch, R, ldst.ptr(), out_pitch, max_label=K)
            torch.cuda.synchronize()
            ldst.check(f"clean R={R}")
            assert mismatches(ldst.data(), want) == 0


def test_bands_peer_halos_prune_fill_blur(ctx, mr, oracle):
    from maprast_b200 import synth
    dev = torch.device("cuda", 0)
    K, R, n1, n2 = 12, 2, 1024, 96
    pal = synth.make_palette(77, K)
    st = synth.make_stats(77, pal)
    ctx.set_palette(st.mean, st.lo, st.hi)
    ctx.palette = None
    ctx.set_known(None)
    scan = oracle.synth_scan(22, 0, pal, n1, n2)
    want = oracle.categorize_clean(scan, st.mean, st.lo, st.hi, R)
    a, b = 32, 64
    body = Guarded(b - a, n1 * 3, n1 * 3, dev, torch.from_numpy(scan[a:b].reshape(b - a, n1 * 3)).to(dev))
    lo = Guarded(R, n1 * 3, n1 * 3, dev, torch.from_numpy(scan[a - R:a].reshape(R, n1 * 3)).to(dev))
    hi = Guarded(R, n1 * 3, n1 * 3, dev, torch.from_numpy(scan[b:b + R].reshape(R, n1 * 3)).to(dev))
    out = Guarded(b - a, n1, n1, dev)
    ctx.categorize_clean_dev_peer(body.ptr(), n1, b - a, n1 * 3, 3, R, lo.ptr(), R, hi.ptr(), R, out.ptr(), n1)
    torch.cuda.synchronize()
    for g, w in ((out, "peer out"), (body, "peer body"), (lo, "peer halo lo"), (hi, "peer halo hi")):
        g.check(w)
    assert mismatches(out.data(), want[a:b]) == 0
    # segment prune and fill_gaps on odd pitches
    n1o, n2o = 333, 77
    rng = np.random.default_rng(4)
    lab = np.kron(rng.integers(1, 6, size=(n2o // 8 + 1, n1o // 8 + 1)), np.ones((8, 8), dtype=np.int64))[:n2o, :n1o].astype(np.uint8)
    lab[rng.random(lab.shape) < 0.08] = 0
    src = Guarded(n2o, 352, n1o, dev, torch.from_numpy(lab).to(dev))
    dst = Guarded(n2o, 368, n1o, dev)
    keep = oracle.keep_table(range(1, 6))
    ctx.clean_segments_dev(src.ptr(), n1o, n2o, 352, keep, 0.01, 10, dst.ptr(), 368)
    torch.cuda.synchronize()
    dst.check("segment prune")
    pruned, _ = oracle.clean_segments(lab, None, range(1, 6), 0.01, 10)
    assert mismatches(dst.data(), pruned) == 0
    px = rng.integers(0, 256, size=(n2o, n1o, 3), dtype=np.uint8)
    mean = rng.random((5, 3))
    ctx.set_palette(mean, mean - 0.1, mean + 0.1)
    psrc = Guarded(n2o, n1o * 3 + 5, n1o * 3, dev, torch.from_numpy(px.reshape(n2o, n1o * 3)).to(dev))
    fdst = Guarded(n2o, 336, n1o, dev)
    ctx.fill_gaps_dev(dst.ptr(), n1o, n2o, 368, psrc.ptr(), n1o * 3 + 5, 3, [1, 2, 3, 4, 5], fdst.ptr(), 336, repeat=3)
    torch.cuda.synchronize()
    fdst.check("fill_gaps")
    dst.check("fill_gaps (input)")
    assert mismatches(fdst.data(), oracle.fill_gaps(pruned, px, mean, [1, 2, 3, 4, 5], repeat=3)) == 0
    # blur into a padded Float64 image
    bdst = Guarded(n2o, n1o * 24 + 40, n1o * 24, dev)
    ctx._ck(ctx._L.mr_blur_dev(ctx._h, psrc.ptr(), n1o, n2o, n1o * 3 + 5, 3, 2, 2, bdst.ptr(), n1o * 24 + 40, None))
    torch.cuda.synchronize()
    bdst.check("blur")
    got = np.frombuffer(bdst.data().tobytes(), dtype=np.float64).reshape(n2o, n1o, 3)
    assert got.tobytes() == oracle.blur(px, 2, 2).tobytes()
    # _balance (per-pixel part) into a padded Float64 image
    coef = np.ascontiguousarray(rng.normal(0, 1e-5, size=(3, 7)))
    coef[:, 0] = rng.normal(0, 0.05, size=3)
    bal = Guarded(n2o, n1o * 24 + 24, n1o * 24, dev)
    ctx._ck(ctx._L.mr_balance_dev(ctx._h, psrc.ptr(), n1o, n2o, n1o * 3 + 5, 3, coef.ctypes.data, bal.ptr(), n1o * 24 + 24, None))
    torch.cuda.synchronize()
    bal.check("balance")
    psrc.check("balance (input)")
    got = np.frombuffer(bal.data().tobytes(), dtype=np.float64).reshape(n2o, n1o, 3)
    assert got.tobytes() == oracle.balance(px, coef).tobytes()
