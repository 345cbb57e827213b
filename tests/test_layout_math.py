"""The tiled bf16 layout is addressed by three pieces of device code that never see each other: the writer
(``tile_db16_kernel``, csrc/build.cu), the GEMM's TMA coordinates (``bc1`` in csrc/gemm_topk.cu, over the 2-D view
``[n_pad * KB][64]`` of csrc/api.cu:make_tmap_tiled) and the batch-1 scan's lane addresses (``scan_scores_tiled_kernel``,
csrc/scan.cu).  Their index arithmetic is restated here and checked against each other on ragged shapes -- a host-side
guard for the one layout every coarse path of a compact index reads."""
import numpy as np
import pytest


def tile_chunk_index(r, kb, c, KB):
    """16-byte chunk index of (row r, 64-column block kb, chunk c) -- tile_db16_kernel's `out`."""
    return ((((r >> 8) * KB + kb) << 8) + (r & 255)) * 8 + c


def tma_row(brow, kb, KB):
    """Row coordinate of a box starting at database row `brow`, k-block kb, in the [n_pad * KB][64] view (gemm_topk.cu)."""
    return (((brow >> 8) * KB + kb) << 8) + (brow & 255)


def scan_lane_chunk(row0, kb, j, lane, KB):
    """Chunk a lane loads in scan_scores_tiled_kernel: base + (kb << 11) + j * 32, base = tile origin + (row0 & 255) * 8 + lane."""
    base = (((row0 >> 8) * KB) << 11) + ((row0 & 255) << 3) + lane
    return base + (kb << 11) + j * 32


@pytest.mark.parametrize("n,d", [(1, 8), (255, 64), (257, 65), (4097, 192), (12345, 520), (70000, 2048)])
def test_writer_tma_and_scan_agree(n, d):
    n_pad, d_pad = (n + 255) // 256 * 256, (d + 63) // 64 * 64
    KB = d_pad // 64
    rng = np.random.default_rng(n + d)
    # the writer is a bijection onto [0, n_pad * d_pad / 8)
    r = np.arange(n_pad, dtype=np.int64)[:, None, None]
    kb = np.arange(KB, dtype=np.int64)[None, :, None]
    c = np.arange(8, dtype=np.int64)[None, None, :]
    out = tile_chunk_index(r, kb, c, KB).ravel()
    assert out.min() == 0 and out.max() == n_pad * KB * 8 - 1 and np.unique(out).size == out.size
    # TMA: a 128- or 256-row box that starts on a multiple of its height is `box` consecutive rows of the 2-D view, each row
    # of the view being one (row, k-block) run of 64 bf16 = 8 chunks
    for box in (128, 256):
        for brow in rng.choice(np.arange(0, n_pad, box), size=min(8, n_pad // box), replace=False):
            for k in rng.integers(0, KB, size=3):
                rows = np.arange(brow, brow + box, dtype=np.int64)
                assert np.array_equal(tma_row(rows, k, KB), tma_row(int(brow), int(k), KB) + np.arange(box))
                assert np.array_equal(tma_row(rows, k, KB) * 8, tile_chunk_index(rows, int(k), 0, KB))
    # scan: lane l of load j in the 16-row group at row0 holds row row0 + 4 j + (l >> 3), chunk l & 7, of k-block kb
    lanes = np.arange(32, dtype=np.int64)
    for row0 in rng.choice(np.arange(0, n_pad, 16), size=min(16, n_pad // 16), replace=False):
        for k in rng.integers(0, KB, size=3):
            for j in range(4):
                want = tile_chunk_index(int(row0) + 4 * j + (lanes >> 3), int(k), lanes & 7, KB)
                assert np.array_equal(scan_lane_chunk(int(row0), int(k), j, lanes, KB), want)
    # and the padding rows the scan's last group and the GEMM's last box read all live in the last 256-row block, which
    # alloc_tiled zeroes as ONE contiguous run at the end of the array (csrc/api.cu)
    if n_pad > n:
        pad = tile_chunk_index(np.arange(n, n_pad, dtype=np.int64)[:, None], np.arange(KB, dtype=np.int64)[None, :], 0, KB)
        assert pad.min() >= (n_pad - 256) * KB * 8
