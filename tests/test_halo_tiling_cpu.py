"""CPU model of the K1h data path (csrc/conv_halo.cu): 10x18 halo tile -> three column-shifted [18][8] copies ->
tap (r, s) = rows [8r, 8r + 128) of copy s, and the `up` variant's 6x10 source box -> 2x2 halo positions.
Pins the index arithmetic the kernel relies on against torch's conv2d (no GPU, no library call)."""
import torch
import torch.nn.functional as Fn

TW, TH = 8, 16          # output pixel box of one CTA
HW, HH = TW + 2, TH + 2


def _copies_from_halo(halo):
    """halo [18][10][C] -> copies [3][18*8][C]: copy s, row y*8 + xx holds halo pixel (y, xx + s)."""
    C = halo.shape[-1]
    copies = torch.zeros(3, HH * 8, C)
    for s in range(3):
        for y in range(HH):
            for xx in range(8):
                copies[s, y * 8 + xx] = halo[y, xx + s]
    return copies


def _tile_conv(copies, w):
    """copies [3][144][C], w [Cout][3][3][C] (KRSC) -> [128][Cout]; output row m = ty*8 + tx."""
    out = torch.zeros(TW * TH, w.shape[0])
    for s in range(3):
        for r in range(3):
            a = copies[s, r * 8: r * 8 + TW * TH]            # the UMMA descriptor copy_s + r * 1024 bytes
            out += a @ w[:, r, s, :].T
    return out


def _run(act, w, halo_fn):
    """act: [H][W][C] activated tensor at the conv's resolution (only used through halo_fn)."""
    H, W, C = act.shape
    out = torch.zeros(H, W, w.shape[0])
    for h0 in range(0, H, TH):
        for w0 in range(0, W, TW):
            halo = halo_fn(h0, w0)
            o = _tile_conv(_copies_from_halo(halo), w)
            out[h0:h0 + TH, w0:w0 + TW] = o.view(TH, TW, -1)
    return out


def test_halo_copies_reproduce_conv3x3_with_zero_padding():
    g = torch.Generator().manual_seed(0)
    H, W, C, Co = 32, 16, 5, 7
    act = torch.randn(H, W, C, generator=g)
    w = torch.randn(Co, 3, 3, C, generator=g)

    def halo_fn(h0, w0):
        halo = torch.zeros(HH, HW, C)                          # out-of-image pixels are 0 AFTER the activation
        for y in range(HH):
            for x in range(HW):
                hh, ww = h0 - 1 + y, w0 - 1 + x
                if 0 <= hh < H and 0 <= ww < W:
                    halo[y, x] = act[hh, ww]
        return halo

    got = _run(act, w, halo_fn)
    want = Fn.conv2d(act.permute(2, 0, 1)[None], w.permute(0, 3, 1, 2), padding=1)[0].permute(1, 2, 0)
    assert torch.allclose(got, want, atol=1e-4)


def test_up_variant_source_box_covers_the_halo():
    """`up` ResBlock: source pixel (lx, ly) of the 6x10 box at ((w0>>1)-1, (h0>>1)-1) lands on halo columns
    {2lx-1, 2lx} and rows {2ly-1, 2ly} (those inside the 10x18 halo)."""
    g = torch.Generator().manual_seed(1)
    Hs, Ws, C, Co = 16, 8, 4, 6
    src = torch.randn(Hs, Ws, C, generator=g)                  # activated half-resolution tensor
    H, W = 2 * Hs, 2 * Ws
    w = torch.randn(Co, 3, 3, C, generator=g)

    def halo_fn(h0, w0):
        halo = torch.full((HH, HW, C), float("nan"))
        for ly in range(10):
            for lx in range(6):
                sh, sw = (h0 >> 1) - 1 + ly, (w0 >> 1) - 1 + lx
                inside = 0 <= sh < Hs and 0 <= sw < Ws
                v = src[sh, sw] if inside else torch.zeros(C)
                for y in (2 * ly - 1, 2 * ly):
                    for x in (2 * lx - 1, 2 * lx):
                        if 0 <= y < HH and 0 <= x < HW:
                            halo[y, x] = v
        assert not torch.isnan(halo).any()                     # every halo pixel is written exactly by the box
        return halo

    up = src.permute(2, 0, 1)[None]
    up = Fn.interpolate(up, scale_factor=2, mode="nearest")
    want = Fn.conv2d(up, w.permute(0, 3, 1, 2), padding=1)[0].permute(1, 2, 0)
    got = _run(torch.zeros(H, W, C), w, halo_fn)
    assert torch.allclose(got, want, atol=1e-4)


def test_transform_thread_mapping_covers_every_vector_once():
    """160 transform threads = (chunk j, halo column x, row parity yh), pixels (x, yh + 2i), i < 9; the `up` variant's
    144 active threads = (j, lx, lyq), source rows lyq + 3i <= 9."""
    seen = set()
    for tt in range(160):
        j, l20 = tt & 7, tt >> 3
        x, yh = l20 % HW, l20 // HW
        for i in range(9):
            seen.add((j, x, yh + 2 * i))
    assert seen == {(j, x, y) for j in range(8) for x in range(HW) for y in range(HH)} and len(seen) == 8 * 180
    seen = []
    for tt in range(160):
        j, l20 = tt & 7, tt >> 3
        lx, lyq = l20 % 6, l20 // 6
        if l20 >= 18:
            continue
        for i in range(4):
            if lyq + 3 * i <= 9:
                seen.append((j, lx, lyq + 3 * i))
    assert len(seen) == len(set(seen)) == 8 * 60
