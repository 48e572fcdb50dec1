"""Pins the tcgen05 shared-memory descriptor semantics conv3d_tc.cu relies on (no-swizzle K-major
core matrices, arbitrary 16-byte-aligned start / leading / stride byte offsets) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)


def _run(n, kblocks, positions, pos0, plane_pad, lbo_positions=None, seed=0):
    """A operand = chunk-planar position array: plane c holds `positions` rows of 8 bf16 (16 B each)."""
    from mvsnet_b200 import ops
    rng = np.random.RandomState(seed)
    nchunks = 2 * kblocks if lbo_positions is None else kblocks
    plane_elems = positions * 8 + plane_pad * 8
    A = rng.randn(nchunks, plane_elems).astype(np.float32)
    A = _bf16(A).float().numpy()
    plane_bytes = plane_elems * 2
    if lbo_positions is None:
        a_kblock_stride, a_lbo = 2 * plane_bytes, plane_bytes        # K block j = chunk planes 2j, 2j+1
    else:
        a_kblock_stride, a_lbo = plane_bytes, lbo_positions * 16     # second K half = same plane shifted
    a_start, a_sbo = pos0 * 16, 128
    # B operand [kblock][khalf][n][8]
    B = _bf16(rng.randn(kblocks, 2, n, 8)).float().numpy()
    b_kblock_stride, b_lbo, b_sbo = 2 * n * 16, n * 16, 128
    out = ops.umma_probe(_bf16(A).cuda(), _bf16(B).cuda(), n, kblocks, a_kblock_stride, a_start, a_lbo, a_sbo,
                         b_kblock_stride, b_lbo, b_sbo).cpu().numpy()
    ref = np.zeros((128, n), dtype=np.float64)
    flat = A.reshape(-1)
    for j in range(kblocks):
        for half in range(2):
            base = (a_start + j * a_kblock_stride + half * a_lbo) // 2
            rows = np.stack([flat[base + m * 8: base + m * 8 + 8] for m in range(128)])      # [128, 8]
            ref += rows.astype(np.float64) @ B[j, half].astype(np.float64).T
    return out, ref


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256])
def test_aligned_tile(n):
    out, ref = _run(n, kblocks=2, positions=128, pos0=0, plane_pad=0)
    assert np.abs(out - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


@pytest.mark.parametrize("pos0", [1, 3, 8, 37])
def test_unaligned_start_and_padded_planes(pos0):
    out, ref = _run(16, kblocks=2, positions=200, pos0=pos0, plane_pad=2)
    assert np.abs(out - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


@pytest.mark.parametrize("lbo_positions", [1, 34, 35])
def test_overlapping_k_halves(lbo_positions):
    """K = 16 built from two 8-channel taps of the same plane (Cin = 8 layers pair taps this way)."""
    out, ref = _run(16, kblocks=3, positions=260, pos0=5, plane_pad=0, lbo_positions=lbo_positions)
    assert np.abs(out - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


def test_many_kblocks():
    out, ref = _run(32, kblocks=8, positions=150, pos0=11, plane_pad=1)
    assert np.abs(out - ref).max() <= 4e-3 * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()
