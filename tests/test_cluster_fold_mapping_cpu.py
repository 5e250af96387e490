"""CPU model of the cluster split-K fold of K1 (csrc/conv_tc.cu: the park loop and cluster_fold_store): the index
arithmetic that decides which CTA / thread finishes which (pixel, channel) of a 128 x BLOCK_N tile, and where a partial
lives in the workspace.  Every output element must be written exactly once, from the S partials in split order."""
import itertools

import numpy as np
import pytest

EPI_THREADS = 256


def park(acc, BLOCK_N):
    """thread = TMEM lane = tile row; warpgroup wg drains column half wg, 32 columns per tcgen05.ld; the partial tile is
    stored [column][row] (wsp + col * 128 + row)."""
    ws = np.full(BLOCK_N * 128, np.nan, dtype=np.float32)
    for wg, row in itertools.product(range(2), range(128)):
        for c in range(wg * (BLOCK_N // 2), (wg + 1) * (BLOCK_N // 2), 32):
            for j in range(32):
                ws[(c + j) * 128 + row] = acc[row, c + j]
    return ws


def fold(parts, BLOCK_N, S, rank, TW, TH, B, n0, writes):
    """cluster_fold_store<BLOCK_N, S> of CTA `rank`: thread g owns rows row..row+3 and channels 8*c8..8*c8+7."""
    k_quads, k_groups = 128 // S // 4, BLOCK_N // 8
    tw_sh, th_sh = TW.bit_length() - 1, TH.bit_length() - 1
    out = {}
    for tid in range(EPI_THREADS):
        for g in range(tid, k_quads * k_groups, EPI_THREADS):
            row, c8 = rank * (128 // S) + (g % k_quads) * 4, g // k_quads
            f = np.zeros((4, 8), dtype=np.float32)
            for s in range(S):                                   # split order, whichever CTA folds
                for j in range(8):
                    f[:, j] = f[:, j] + parts[s][(c8 * 8 + j) * 128 + row: (c8 * 8 + j) * 128 + row + 4]
            wl, hl, nl = row & (TW - 1), (row >> tw_sh) & (TH - 1), row >> (tw_sh + th_sh)
            if n0 + nl >= B:
                continue
            for q in range(4):
                # the quad must stay inside one box row: pixel (wl + q, hl, nl)
                assert wl + q < TW
                for j in range(8):
                    key = (nl, hl, wl + q, c8 * 8 + j)
                    assert key not in writes
                    writes[key] = f[q, j]
    return out


@pytest.mark.parametrize("BLOCK_N,S", [(64, 2), (64, 4), (64, 8), (128, 2), (128, 4), (128, 8), (256, 2), (256, 4), (256, 8)])
@pytest.mark.parametrize("TW,TH,TN,B", [(8, 8, 2, 1), (8, 8, 2, 2), (16, 8, 1, 1), (4, 4, 8, 5), (128, 1, 1, 1)])
def test_cluster_fold_covers_the_tile_once_in_split_order(BLOCK_N, S, TW, TH, TN, B):
    assert TW * TH * TN == 128
    rng = np.random.RandomState(BLOCK_N + S + TW)
    accs = [rng.standard_normal((128, BLOCK_N)).astype(np.float32) for _ in range(S)]
    parts = [park(a, BLOCK_N) for a in accs]
    assert not any(np.isnan(p).any() for p in parts)              # the park loop fills the whole partial tile
    writes = {}
    for rank in range(S):
        fold(parts, BLOCK_N, S, rank, TW, TH, B, 0, writes)
    valid_rows = [r for r in range(128) if r // (TW * TH) < B]
    assert len(writes) == len(valid_rows) * BLOCK_N
    want = np.zeros((128, BLOCK_N), dtype=np.float32)
    for a in accs:                                                # the same left-to-right fp32 sum
        want = want + a
    for r in valid_rows:
        wl, hl, nl = r % TW, (r // TW) % TH, r // (TW * TH)
        for c in range(0, BLOCK_N, 37):
            assert writes[(nl, hl, wl, c)] == want[r, c]
