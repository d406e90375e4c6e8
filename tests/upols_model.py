"""numpy model of the DEVICE algorithm (uniform-partitioned overlap-add driven by the C++ planner's output).

Test infrastructure: it mirrors what k_ir_fft / k_x_fft / k_cmac / k_ifft_ola do, block for block, in float64, so the
host planner (alr_debug_plan, no GPU needed) can be validated on CPU against the oracle.
"""
import numpy as np


def crossfade_gain(plan_ir, wband, t):
    """g_l(t) exactly as k_x_fft evaluates it from the weight band."""
    xb0, xnb, xslot, woff, jmin, nrows = [int(v) for v in plan_ir]
    q = (t >> 7) - jmin
    p = t & 127
    s = np.sin(np.pi * p / 256.0) ** 2
    w0 = np.where((q >= 0) & (q < nrows), wband[np.clip(woff + q, 0, len(wband) - 1)], 0.0)
    w1 = np.where((q + 1 >= 0) & (q + 1 < nrows), wband[np.clip(woff + q + 1, 0, len(wband) - 1)], 0.0)
    return w0 * (1 - s) + w1 * s


def model_convolve(audio, irs, plan, scales, moving, n_out):
    """audio (Lx,), irs (C, N, Lh), scales (N,) -> (C, n_out) following the plan."""
    P = plan["P"]
    C, N, Lh = irs.shape
    K, B_valid, n_valid, xlimit = plan["K"], plan["B_valid"], plan["n_valid"], plan["xlimit"]
    x = np.zeros(max(xlimit, 1) + 2 * P)
    x[:xlimit] = audio[:xlimit]
    H = np.zeros((N, K, C, P + 1), dtype=complex)
    for l in range(N):
        for k in range(K):
            seg = np.zeros((C, 2 * P))
            part = irs[:, l, k * P:(k + 1) * P]
            seg[:, :part.shape[1]] = part
            H[l, k] = np.fft.rfft(seg, axis=-1)
    X = {}
    for l in range(N):
        xb0, xnb = int(plan["irs"][l][0]), int(plan["irs"][l][1])
        for j in range(xnb):
            t = (xb0 + j) * P + np.arange(P)
            blk = x[t] * scales[l]
            if moving:
                blk = blk * crossfade_gain(plan["irs"][l], plan["wband"], t)
            blk = np.where(t < xlimit, blk, 0.0)
            X[(l, j)] = np.fft.rfft(np.concatenate([blk, np.zeros(P)]))
    out = np.zeros((C, (B_valid + 2) * P))
    for b in range(B_valid):
        Y = np.zeros((C, P + 1), dtype=complex)
        lmin, lmax = plan["lrange"][b]
        for l in range(lmin, lmax + 1):
            xb0, xnb = int(plan["irs"][l][0]), int(plan["irs"][l][1])
            d = b - xb0
            for j in range(max(0, d - K + 1), min(xnb - 1, d) + 1):
                Y += X[(l, j)][None, :] * H[l, d - j]
        out[:, b * P:(b + 2) * P] += np.fft.irfft(Y, 2 * P, axis=-1)
    res = np.zeros((C, n_out))
    res[:, :n_valid] = out[:, :n_valid]
    return res
