"""CPU oracles for the algebra behind ``ops/csrc/conv1_fused.cu`` (no GPU, plain PyTorch).

The fused first VBM block rests on three identities that are easy to get wrong in index arithmetic, so each one is
re-derived here with dense tensors and checked against ``torch.nn.functional`` / autograd:

1. a C_in = 1 3x3x3 convolution is a sum over (kd, kh) of ``X_shifted @ T_{kd,kh}`` with banded Toeplitz matrices
   over the W axis (the GEMM the tensor cores run, 8 output columns x 16 channels per block, K = 16 input columns);
2. BatchNorm -> ReLU -> MaxPool(2) forward as ``pooled value + code byte`` (arg-max position d*4 + h*2 + w, bit 3 =
   ReLU active), and its backward written as ``dy = S*g + (A + B*y)`` with per-channel constants;
3. the weight gradient as ``dW~[kh][(wl, c), kd*16 + k] = sum_rows dy[row, (wl, c)] * XP[line h+kh, row+kd, 8j+k]``
   accumulated over all blocks j, followed by the diagonal extraction ``dW[c,kd,kh,kw] = sum_wl dW~[kh][(wl,c), kd*16+wl+kw]``.
"""
import torch
import torch.nn.functional as F


def _pad_hd(x):
    """[N,D,H,W] -> XP2[(n, h', d'), w'] with zero halo, W padded to 8*(ceil(W/8)+1) (conv1_pad_input_hd_kernel)."""
    N, D, H, W = x.shape
    nblk = (W + 7) // 8
    Wq = 8 * (nblk + 1)
    xp = torch.zeros(N, H + 2, D + 2, Wq, dtype=x.dtype)
    xp[:, 1:H + 1, 1:D + 1, 1:W + 1] = x.permute(0, 2, 1, 3)
    return xp, nblk


def _toeplitz(w):
    """w: [16,1,3,3,3] -> T[kd][kh]: [16 (k), 128 (wl*16 + c)] with T[k, wl*16+c] = w[c, kd, kh, k - wl]."""
    T = torch.zeros(3, 3, 16, 128, dtype=w.dtype)
    for wl in range(8):
        for kw in range(3):
            T[:, :, wl + kw, wl * 16:(wl + 1) * 16] = w[:, 0, :, :, kw].permute(1, 2, 0)
    return T


def _conv_by_toeplitz(x, w):
    """Y[n,d,h,w,c] through the per-block GEMMs of the kernel."""
    N, D, H, W = x.shape
    xp, nblk = _pad_hd(x)
    T = _toeplitz(w)
    y = torch.zeros(N, D, H, nblk * 8, 16, dtype=x.dtype)
    for j in range(nblk):
        for kd in range(3):
            for kh in range(3):
                # operand of (kd, kh): lines h + kh, rows d + kd, window columns 8j .. 8j+15
                a = xp[:, kh:kh + H, kd:kd + D, 8 * j:8 * j + 16]                     # [N, H, D, 16]
                blk = (a.reshape(-1, 16) @ T[kd, kh]).reshape(N, H, D, 8, 16)
                y[:, :, :, 8 * j:8 * j + 8] += blk.permute(0, 2, 1, 3, 4)
    return y[:, :, :, :W]


def test_toeplitz_gemm_equals_conv3d():
    torch.manual_seed(0)
    for shape in ((2, 5, 6, 13), (1, 4, 3, 24), (1, 3, 3, 8)):
        x = torch.randn(*shape, dtype=torch.float64)
        w = torch.randn(16, 1, 3, 3, 3, dtype=torch.float64)
        ref = F.conv3d(x.unsqueeze(1), w, padding=1).permute(0, 2, 3, 4, 1)
        assert torch.allclose(_conv_by_toeplitz(x, w), ref, atol=1e-10)


def _pool_with_code(z):
    """z: [N,D,H,W,C] (post BN, pre ReLU) -> pooled relu-max [N,D/2,H/2,W/2,C], code = argpos | 8*active (kernel rule)."""
    N, D, H, W, C = z.shape
    PD, PH, PW = D // 2, H // 2, W // 2
    win = z[:, :2 * PD, :2 * PH, :2 * PW].reshape(N, PD, 2, PH, 2, PW, 2, C).permute(0, 1, 3, 5, 7, 2, 4, 6).reshape(N, PD, PH, PW, C, 8)
    m, k = win.max(-1)                                           # position index = d*4 + h*2 + w
    active = m > 0
    return torch.where(active, m, torch.zeros_like(m)), k + 8 * active.long()


def test_pool_code_and_folded_bn_backward_match_autograd():
    torch.manual_seed(1)
    N, D, H, W = 2, 5, 6, 7                                       # odd sizes: edge voxels belong to no pooling window
    x = torch.randn(N, D, H, W, dtype=torch.float64)
    w = torch.randn(16, 1, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(16, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.randn(16, dtype=torch.float64) * 0.2).requires_grad_(True)
    eps = 1e-5
    y = F.conv3d(x.unsqueeze(1), w, padding=1)
    mean, var = y.mean((0, 2, 3, 4)), y.var((0, 2, 3, 4), unbiased=False)
    istd = (var + eps).rsqrt()
    bc = lambda t: t[None, :, None, None, None]
    z = (y - bc(mean)) * bc(istd) * bc(gamma) + bc(beta)
    p_ref = F.max_pool3d(torch.relu(z), 2)
    dp = torch.randn_like(p_ref)
    p_ref.backward(dp)

    # ---- forward the kernel's way
    yl = y.detach().permute(0, 2, 3, 4, 1)                         # channels last
    sc = gamma.detach() * istd.detach()
    sh = beta.detach() - mean.detach() * sc
    p, code = _pool_with_code(yl * sc + sh)
    assert torch.allclose(p, p_ref.detach().permute(0, 2, 3, 4, 1), atol=1e-12)

    # ---- backward the kernel's way: sums over the pooled tensors, then dy = S*g + (A + B*y) for EVERY voxel
    dpl = dp.permute(0, 2, 3, 4, 1)
    active = code >= 8
    g_pooled = torch.where(active, dpl, torch.zeros_like(dpl))
    xhat_at_max = (p - beta.detach()) / gamma.detach()              # recovered from the pooled output (pooled-stats kernel)
    count = N * D * H * W
    c1 = g_pooled.sum((0, 1, 2, 3)) / count                         # mean of g        -> dbeta / count
    c2 = (g_pooled * xhat_at_max).sum((0, 1, 2, 3)) / count         # mean of g * xhat -> dgamma / count
    assert torch.allclose(c1 * count, beta.grad, atol=1e-9) and torch.allclose(c2 * count, gamma.grad, atol=1e-9)
    S = sc
    B = -S * c2 * istd.detach()
    A = -S * c1 - B * mean.detach()
    g_full = torch.zeros_like(yl)
    PD, PH, PW = D // 2, H // 2, W // 2
    for dd in range(2):
        for hh in range(2):
            for ww in range(2):
                hit = (code == (dd * 4 + hh * 2 + ww + 8))
                g_full[:, dd:2 * PD:2, hh:2 * PH:2, ww:2 * PW:2] = torch.where(hit, dpl, torch.zeros_like(dpl))
    dy = S * g_full + A + B * yl
    # autograd's gradient wrt the conv output
    y2 = y.detach().clone().requires_grad_(True)
    m2, v2 = y2.mean((0, 2, 3, 4)), y2.var((0, 2, 3, 4), unbiased=False)
    z2 = (y2 - bc(m2)) * bc((v2 + eps).rsqrt()) * bc(gamma.detach()) + bc(beta.detach())
    F.max_pool3d(torch.relu(z2), 2).backward(dp)
    assert torch.allclose(dy, y2.grad.permute(0, 2, 3, 4, 1), atol=1e-9)


def test_shifted_window_wgrad_with_diagonal_extraction():
    """dW1 from dy and the padded input exactly as the second GEMM of the backward kernel produces it."""
    torch.manual_seed(2)
    N, D, H, W = 2, 4, 5, 11
    x = torch.randn(N, D, H, W, dtype=torch.float64)
    dy = torch.randn(N, D, H, W, 16, dtype=torch.float64)           # any gradient of the conv output, channels last
    w = torch.zeros(16, 1, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv3d(x.unsqueeze(1), w, padding=1).backward(dy.permute(0, 4, 1, 2, 3))
    xp, nblk = _pad_hd(x)
    dyp = torch.zeros(N, D, H, nblk * 8, 16, dtype=torch.float64)
    dyp[:, :, :, :W] = dy                                           # columns beyond W are written as zeros by the epilogue
    acc = torch.zeros(3, 128, 48, dtype=torch.float64)              # dW~[kh][(wl, c), kd*16 + k]
    for j in range(nblk):
        a = dyp[:, :, :, 8 * j:8 * j + 8].permute(0, 2, 1, 3, 4).reshape(N, H, D, 128)       # rows (n, h, d), M = (wl, c)
        for kh in range(3):
            for kd in range(3):                                     # the three N blocks = row-shifted views of one window
                b = xp[:, kh:kh + H, kd:kd + D, 8 * j:8 * j + 16]                            # [N, H, D, 16]
                acc[kh, :, kd * 16:(kd + 1) * 16] += a.reshape(-1, 128).t() @ b.reshape(-1, 16)
    dw = torch.zeros(16, 3, 3, 3, dtype=torch.float64)
    for wl in range(8):
        for kd in range(3):
            for kh in range(3):
                for kw in range(3):
                    dw[:, kd, kh, kw] += acc[kh, wl * 16:(wl + 1) * 16, kd * 16 + wl + kw]
    assert torch.allclose(dw, w.grad[:, 0], atol=1e-9)
