"""CPU oracles for the addressing tricks of the tcgen05 convolution kernels (plain PyTorch, no GPU).

* ``conv3d_halo.cu``: a tile is TH rows of one (n, d) plane in the zero-PADDED, flattened pixel order (width Wp = W+2);
  the operand of tap (kd, kh, kw) is the same halo buffer shifted by ``kh*Wp + kw`` pixels, outputs that fall on the two
  padding columns are computed and dropped.
* ``conv3d_wgrad_halo.cu``: for a filter row (kd, kh) the three kw taps are pixel-shifted views of one halo plane
  (M blocks one pixel apart), so one MMA per 16 pixels yields dW for kw = 0..2 (further shifts are discarded).
* ``conv3d_wgrad_tap.cu``: per tap, ``dW_tap = X_box(shifted, zero outside the volume)^T @ dY_box`` over voxel boxes.
"""
import torch
import torch.nn.functional as F


def test_halo_shifted_view_convolution():
    torch.manual_seed(0)
    N, D, H, W, CI, CO, TH = 1, 4, 5, 6, 3, 4, 2
    x = torch.randn(N, D, H, W, CI, dtype=torch.float64)
    w = torch.randn(CO, CI, 3, 3, 3, dtype=torch.float64)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=1).permute(0, 2, 3, 4, 1)
    Wp = W + 2
    xpad = F.pad(x, (0, 0, 1, 1, 1, 1, 1, 1))                       # zero halo in d, h, w (what TMA OOB fill produces)
    y = torch.zeros(N, D, H, W, CO, dtype=torch.float64)
    for d in range(D):
        for h0 in range(0, H, TH):
            rows = min(TH, H - h0)
            # three halo planes (kd), each (TH+2) x Wp pixels, flattened; + slack for the shifted reads
            planes = [torch.cat([xpad[0, d + kd, h0:h0 + TH + 2].reshape(-1, CI), torch.zeros(2 * Wp, CI, dtype=torch.float64)])
                      for kd in range(3)]
            acc = torch.zeros(TH * Wp, CO, dtype=torch.float64)     # one accumulator row per PADDED pixel of the tile
            for kd in range(3):
                for kh in range(3):
                    for kw in range(3):
                        shift = kh * Wp + kw
                        acc += planes[kd][shift:shift + TH * Wp] @ w[:, :, kd, kh, kw].t()
            tile = acc.reshape(TH, Wp, CO)[:rows, :W]               # drop the two padding columns of every line
            y[0, d, h0:h0 + rows] = tile
    assert torch.allclose(y, ref, atol=1e-10)


def test_halo_wgrad_pixel_shifted_blocks():
    torch.manual_seed(1)
    N, D, H, W, CI, CO = 1, 3, 4, 5, 2, 3
    x = torch.randn(N, D, H, W, CI, dtype=torch.float64)
    dy = torch.randn(N, D, H, W, CO, dtype=torch.float64)
    w = torch.zeros(CO, CI, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=1).backward(dy.permute(0, 4, 1, 2, 3))
    Wp = W + 2
    xpad = F.pad(x, (0, 0, 1, 1, 1, 1, 1, 1))
    dyp = F.pad(dy, (0, 0, 0, 2))                                   # dy box of width Wp: the two extra columns arrive as zeros
    dw = torch.zeros(3, 3, 3, CI, CO, dtype=torch.float64)
    for d in range(D):
        dyt = dyp[0, d].reshape(-1, CO)                             # [H*Wp, CO], padded pixel order
        for kd in range(3):
            plane = torch.cat([xpad[0, d + kd].reshape(-1, CI), torch.zeros(8, CI, dtype=torch.float64)])
            for kh in range(3):
                base = kh * Wp
                for j in range(8):                                  # M blocks = views shifted by j pixels; only j < 3 is kept
                    blk = plane[base + j:base + j + H * Wp].t() @ dyt
                    if j < 3:
                        dw[kd, kh, j] += blk
    assert torch.allclose(dw.permute(4, 3, 0, 1, 2), w.grad, atol=1e-10)


def test_per_tap_box_wgrad():
    torch.manual_seed(2)
    N, D, H, W, CI, CO = 2, 3, 5, 4, 3, 2
    Hb, Db = 2, 2                                                   # voxel box = W x Hb x Db; ragged at the volume end
    x = torch.randn(N, D, H, W, CI, dtype=torch.float64)
    dy = torch.randn(N, D, H, W, CO, dtype=torch.float64)
    w = torch.zeros(CO, CI, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=1).backward(dy.permute(0, 4, 1, 2, 3))

    def box(t, n, d0, h0, w0):
        """TMA box {C, W, Hb, Db} at (w0, h0, d0): out-of-bounds voxels are zero-filled."""
        out = torch.zeros(Db, Hb, W, t.shape[-1], dtype=torch.float64)
        for dd in range(Db):
            for hh in range(Hb):
                for ww in range(W):
                    d, h, wv = d0 + dd, h0 + hh, w0 + ww
                    if 0 <= d < D and 0 <= h < H and 0 <= wv < W:
                        out[dd, hh, ww] = t[n, d, h, wv]
        return out.reshape(-1, t.shape[-1])

    dw = torch.zeros(3, 3, 3, CI, CO, dtype=torch.float64)
    for n in range(N):
        for d0 in range(0, D, Db):
            for h0 in range(0, H, Hb):
                dyb = box(dy, n, d0, h0, 0)
                for kd in range(3):
                    for kh in range(3):
                        for kw in range(3):
                            dw[kd, kh, kw] += box(x, n, d0 + kd - 1, h0 + kh - 1, kw - 1).t() @ dyb
    assert torch.allclose(dw.permute(4, 3, 0, 1, 2), w.grad, atol=1e-10)
