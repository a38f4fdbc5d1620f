"""CPU emulation of the shared-memory layout / fragment indexing of csrc/conv_c32.cu (the experimental mma.sync 3x3 conv for
the 32-channel HRNet branch).  The kernel could not be run on a B200 when it was written, so its index arithmetic -- swizzled
64-byte rows, tap-shifted ldmatrix addresses, m16n8k16 fragment ownership, output staging -- is transcribed here line by line
and executed with the documented PTX semantics of ldmatrix (.x4, non-transposed) and mma.m16n8k16 (row.col):

  ldmatrix: lanes 8i..8i+7 supply the 8 row addresses (16 bytes each) of matrix i; lane l receives from matrix i the two
            elements (row l>>2, columns 2*(l&3), 2*(l&3)+1) in register i.
  mma A   : a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 (row g, k+8), a3 (row g+8, k+8);  g = l>>2, t = l&3
  mma B   : b0 (k 2t..2t+1, col g), b1 (k 2t+8.., col g)
  mma C/D : c0,c1 (row g, cols 2t, 2t+1), c2,c3 (row g+8, same cols)

The result must equal a direct convolution with zero padding.  This checks the indexing, not the hardware."""
import numpy as np

C, TH, TW = 32, 4, 32
HH, HW = TH + 2, TW + 2


def off(row, c):                       # c32_off: byte offset of 16-byte chunk c of 64-byte row `row`
    return row * 64 + ((c ^ ((row >> 1) & 3)) << 4)


def ldsm_x4(mem, addrs):
    """mem: element array (2-byte elements); addrs[lane] = byte address.  Returns regs[lane][i] = (elem, elem) pairs."""
    regs = [[None] * 4 for _ in range(32)]
    for i in range(4):
        rows = [mem[addrs[8 * i + r] // 2: addrs[8 * i + r] // 2 + 8] for r in range(8)]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            regs[lane][i] = (rows[g][2 * t], rows[g][2 * t + 1])
    return regs


def mma(acc, a, b0, b1):
    """acc[lane][4] += A(16x16) . B(16x8) with the fragment ownership listed in the module docstring"""
    A = np.zeros((16, 16)); Bm = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = a[lane][0]
        A[g + 8, 2 * t:2 * t + 2] = a[lane][1]
        A[g, 2 * t + 8:2 * t + 10] = a[lane][2]
        A[g + 8, 2 * t + 8:2 * t + 10] = a[lane][3]
        Bm[2 * t:2 * t + 2, g] = b0[lane]
        Bm[2 * t + 8:2 * t + 10, g] = b1[lane]
    D = A @ Bm
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += D[g, 2 * t]; acc[lane][1] += D[g, 2 * t + 1]
        acc[lane][2] += D[g + 8, 2 * t]; acc[lane][3] += D[g + 8, 2 * t + 1]


def emulate_tile(x, wp, dys, dxs, b, y0, x0, H, W, y):
    in_sm = np.zeros(HH * HW * 32)
    w_sm = np.zeros(9 * C * 32)
    for i in range(9 * C * 4):                                   # weight staging loop of the kernel
        q, c = i >> 2, i & 3
        w_sm[off(q, c) // 2: off(q, c) // 2 + 8] = wp.reshape(-1)[q * C + c * 8: q * C + c * 8 + 8]
    for i in range(HH * HW * 4):                                 # halo tile
        p, c = i >> 2, i & 3
        ry, rx = divmod(p, HW)
        gy, gx = y0 - 1 + ry, x0 - 1 + rx
        inside = 0 <= gy < H and 0 <= gx < W
        in_sm[off(p, c) // 2: off(p, c) // 2 + 8] = x[b, gy, gx, c * 8:c * 8 + 8] if inside else 0.0
    for warp in range(TH):
        acc = [[[[0.0] * 4 for _ in range(32)] for _ in range(4)] for _ in range(2)]      # [mt][nt][lane][4]
        for t in range(9):
            prow = (warp + 1 + dys[t]) * HW + 1 + dxs[t]
            for ks in range(2):
                a = [ldsm_x4(in_sm, [off(prow + mt * 16 + (lane & 15), ks * 2 + (lane >> 4)) for lane in range(32)]) for mt in range(2)]
                bw = [ldsm_x4(w_sm, [off(t * C + (np_ * 2 + (lane >> 4)) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1))
                                     for lane in range(32)]) for np_ in range(2)]
                for mt in range(2):
                    for nt in range(4):
                        mma(acc[mt][nt], a[mt],
                            [bw[nt >> 1][lane][(nt & 1) * 2] for lane in range(32)],
                            [bw[nt >> 1][lane][(nt & 1) * 2 + 1] for lane in range(32)])
        # output staging (warp-private 2 KB of the old halo tile) and the 16-byte stores
        stage = np.zeros(HH * HW * 32)
        wbase = warp * (TW * 64)
        for lane in range(32):
            g, tq = lane >> 2, lane & 3
            for mt in range(2):
                for half in range(2):
                    px = mt * 16 + g + half * 8
                    for nt in range(4):
                        e = (wbase + off(px, nt) + tq * 4) // 2
                        stage[e], stage[e + 1] = acc[mt][nt][lane][half * 2], acc[mt][nt][lane][half * 2 + 1]
        oy = y0 + warp
        if oy < H:
            for lane in range(32):
                for j in range(4):
                    px, c = j * 8 + (lane >> 2), lane & 3
                    ox = x0 + px
                    if ox < W:
                        e = (wbase + off(px, c)) // 2
                        y[b, oy, ox, c * 8:c * 8 + 8] = stage[e:e + 8]


def reference(x, wp, dys, dxs):
    B, H, W, _ = x.shape
    y = np.zeros((B, H, W, C))
    xp = np.zeros((B, H + 2, W + 2, C)); xp[:, 1:-1, 1:-1] = x
    for t in range(9):
        y += np.einsum("bhwk,nk->bhwn", xp[:, 1 + dys[t]:1 + dys[t] + H, 1 + dxs[t]:1 + dxs[t] + W], wp[t])
    return y


def test_conv_c32_indexing_matches_direct_convolution():
    rng = np.random.default_rng(5)
    B, H, W = 1, 6, 37                                  # ragged: second tile row and second tile column are partial
    x = rng.standard_normal((B, H, W, C))
    wp = rng.standard_normal((9, C, C))                 # packed [tap][n][k]
    for negate in (False, True):                        # forward taps and the negated taps of the data-gradient pack
        dys = [(t // 3 - 1) * (-1 if negate else 1) for t in range(9)]
        dxs = [(t % 3 - 1) * (-1 if negate else 1) for t in range(9)]
        y = np.full((B, H, W, C), np.nan)
        for ty in range((H + TH - 1) // TH):
            for tx in range((W + TW - 1) // TW):
                emulate_tile(x, wp, dys, dxs, 0, ty * TH, tx * TW, H, W, y)
        ref = reference(x, wp, dys, dxs)
        assert not np.isnan(y).any()
        assert np.abs(y - ref).max() < 1e-9


def test_swizzle_is_conflict_free_for_ldmatrix_and_staging():
    # the 8 row addresses of one ldmatrix matrix (8 consecutive rows, same logical chunk) must fall in 8 distinct 16-byte bank groups
    for start in range(0, 64):
        for c in range(4):
            groups = {(off(start + r, c) % 128) // 16 for r in range(8)}
            assert len(groups) == 8
    # staging writes: 32 lanes (g = 0..7 pixels, tq = 0..3 words) of one (mt, half, nt) store -> 32 distinct banks
    for nt in range(4):
        banks = {((off(g, nt) + tq * 4) % 128) // 4 for g in range(8) for tq in range(4)}
        assert len(banks) == 32
