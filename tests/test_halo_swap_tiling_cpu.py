"""CPU model of the K1s data path (csrc/conv_halo_swap.cu): 18x18 halo tile -> three column-shifted [18][16] copies ->
tap (r, s) = 256 consecutive pixel rows starting at row 16 r of copy s (the B operand), weights = A operand, accumulator
lane = output channel / column = pixel, epilogue = (warpgroup, chunk, q) -> pixel; plus the integer bookkeeping the kernel
relies on (thread mapping, rotated coefficient table, ring-stage alignment of the skip source).  No GPU, no library call."""
import torch
import torch.nn.functional as Fn

T, HALO = 16, 18


def _copies(halo):
    """halo [18][18][C] -> copies [3][18*16][C]: copy s, row y*16 + xx holds halo pixel (y, xx + s)."""
    C = halo.shape[-1]
    c = torch.zeros(3, HALO * T, C)
    for s in range(3):
        for y in range(HALO):
            for xx in range(T):
                c[s, y * T + xx] = halo[y, xx + s]
    return c


def test_swapped_roles_reproduce_conv3x3():
    g = torch.Generator().manual_seed(0)
    H, W, C, Co = 32, 48, 5, 7
    act = torch.randn(H, W, C, generator=g)
    w = torch.randn(Co, 3, 3, C, generator=g)                   # KRSC
    out = torch.zeros(H, W, Co)
    for h0 in range(0, H, T):
        for w0 in range(0, W, T):
            halo = torch.zeros(HALO, HALO, C)                    # out-of-image pixels are 0 AFTER the activation
            for y in range(HALO):
                for x in range(HALO):
                    hh, ww = h0 - 1 + y, w0 - 1 + x
                    if 0 <= hh < H and 0 <= ww < W:
                        halo[y, x] = act[hh, ww]
            cp = _copies(halo)
            d = torch.zeros(Co, T * T)                           # D^T: lane = output channel, column = pixel
            for s in range(3):
                for r in range(3):
                    b = cp[s, r * T: r * T + T * T]              # descriptor copy_s + r * 2048 bytes: 256 pixel rows
                    d += w[:, r, s, :] @ b.T                     # A = weights [Co][C], B = pixels [256][C]
            # epilogue: warpgroup wg drains columns [128 wg, 128 wg + 128): chunk = box row 8 wg + chunk, q = box column
            for wg in range(2):
                for chunk in range(8):
                    for q in range(16):
                        col = wg * 128 + chunk * 16 + q
                        out[h0 + wg * 8 + chunk, w0 + q] = d[:, col]
    want = Fn.conv2d(act.permute(2, 0, 1)[None], w.permute(0, 3, 1, 2), padding=1)[0].permute(1, 2, 0)
    assert torch.allclose(out, want, atol=1e-4)


def test_transform_thread_mapping_and_swizzle():
    """288 transform threads = (16-byte chunk j, halo column x, half yh), pixels (x, 9 yh + i), i < 9: every (chunk, pixel)
    once; the raw tile's 128-byte swizzle term of pixel px = y*18 + x is (px & 7); copy rows are 128-byte swizzled by
    (xx & 7) inside 8-pixel atoms and never collide."""
    seen = set()
    for tt in range(288):
        j, l36 = tt & 7, tt >> 3
        x, yh = l36 % HALO, l36 // HALO
        px0 = yh * 9 * HALO + x
        for i in range(9):
            y = yh * 9 + i
            assert (px0 + i * HALO) == y * HALO + x
            seen.add((j, x, y))
    assert seen == {(j, x, y) for j in range(8) for x in range(HALO) for y in range(HALO)}
    addrs = set()
    for s in range(3):
        for y in range(HALO):
            for x in range(HALO):
                xx = x - s
                if not 0 <= xx < T:
                    continue
                for j in range(8):
                    a = s * HALO * 2048 + y * 2048 + ((xx >> 3) & 1) * 1024 + (xx & 7) * 128 + ((j ^ (xx & 7)) << 4)
                    assert a not in addrs
                    addrs.add(a)
    assert len(addrs) == 3 * HALO * T * 8


def test_coefficient_table_rotation_is_a_bijection_and_conflict_free():
    """float4 i (channels 2i, 2i+1) of the per-image coefficient table is stored at chunk (i >> 2), quarter
    ((i & 3) + ((chunk & 7) >> 1)) & 3; a warp's load of quarter q (8 chunks j, each read by 4 lanes) must touch 8
    different 16-byte bank groups."""
    C = 256
    off = {}
    for i in range(C // 2):
        ch = i >> 2
        o = ch * 64 + (((i & 3) + ((ch & 7) >> 1)) & 3) * 16
        assert o not in off.values()
        off[i] = o
    assert sorted(off.values()) == list(range(0, C * 8, 16))
    for kc in range(C // 64):
        for q in range(4):
            groups = set()
            for j in range(8):
                a = kc * 512 + j * 64 + ((q + (j >> 1)) & 3) * 16        # the reader's address
                assert a == off[(kc * 8 + j) * 4 + q]                     # reads what the writer stored
                groups.add((a // 16) % 8)
            assert len(groups) == 8


def test_skip_source_ring_alignment():
    """The two 16x8 pixel boxes of a skip-source slice must sit in adjacent ring stages (one 256-row operand): with 4 stages
    and the rule "skip the last stage before a pair", producer and consumer stay in lock step for every kc1 / kc2."""
    stages = 4
    for kc1 in range(1, 17):
        for kc2 in range(0, 9):
            stage, log = 0, []
            for tile in range(5):
                stage = (stage + 9 * kc1) % stages
                for _ in range(kc2):
                    if stage == stages - 1:
                        stage = (stage + 1) % stages
                    assert stage + 1 < stages                       # the pair does not wrap
                    log.append(stage)
                    stage = (stage + 3) % stages
            assert all(s in (0, 1, 2) for s in log)


def test_up_variant_source_box_covers_the_halo():
    """K1s `up` variant: the 18x18 halo of the upsampled image is covered exactly once by the 10x10 source box at
    ((w0>>1)-1, (h0>>1)-1): source pixel (lx, ly) -> halo columns {2lx-1, 2lx}, rows {2ly-1, 2ly}; 240 of the 288 transform
    threads = (chunk j, lx, lyq) own source rows lyq + 3i < 10."""
    g = torch.Generator().manual_seed(1)
    Hs, Ws, C, Co = 16, 24, 3, 4
    src = torch.randn(Hs, Ws, C, generator=g)
    H, W = 2 * Hs, 2 * Ws
    w = torch.randn(Co, 3, 3, C, generator=g)
    out = torch.zeros(H, W, Co)
    for h0 in range(0, H, T):
        for w0 in range(0, W, T):
            halo = torch.full((HALO, HALO, C), float("nan"))
            for ly in range(10):
                for lx in range(10):
                    sh, sw = (h0 >> 1) - 1 + ly, (w0 >> 1) - 1 + lx
                    v = src[sh, sw] if (0 <= sh < Hs and 0 <= sw < Ws) else torch.zeros(C)
                    for y in (2 * ly - 1, 2 * ly):
                        for x in (2 * lx - 1, 2 * lx):
                            if 0 <= y < HALO and 0 <= x < HALO:
                                assert torch.isnan(halo[y, x]).all()          # written exactly once
                                halo[y, x] = v
            assert not torch.isnan(halo).any()
            cp = _copies(halo)
            d = torch.zeros(Co, T * T)
            for s in range(3):
                for r in range(3):
                    d += w[:, r, s, :] @ cp[s, r * T: r * T + T * T].T
            out[h0:h0 + T, w0:w0 + T] = d.T.reshape(T, T, Co)
    up = Fn.interpolate(src.permute(2, 0, 1)[None], scale_factor=2, mode="nearest")
    want = Fn.conv2d(up, w.permute(0, 3, 1, 2), padding=1)[0].permute(1, 2, 0)
    assert torch.allclose(out, want, atol=1e-4)
    seen = []
    for tt in range(288):
        j, l36 = tt & 7, tt >> 3
        lx, lyq = l36 % 10, l36 // 10
        if l36 >= 30:
            continue
        for i in range(4):
            if lyq + 3 * i < 10:
                seen.append((j, lx, lyq + 3 * i))
    assert len(seen) == len(set(seen)) == 8 * 100
