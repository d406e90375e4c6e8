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


def kernel_constants():
    """kGm (ALR_CMAC_G), kCmacRuns, kMaxHeads, kMaxItems as compiled into k_cmac (read from alr_kernels.cuh so that the
    model cannot drift from the kernel unnoticed)."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audiblelight_b200", "csrc",
                            "alr_kernels.cuh")).read()
    return dict(G=int(re.search(r"#define ALR_CMAC_G (\d+)", src).group(1)),
                runs=int(re.search(r"#define ALR_CMAC_RUNS (\d+)", src).group(1)),
                max_heads=int(re.search(r"constexpr int kMaxHeads = (\d+);", src).group(1)),
                max_items=int(re.search(r"constexpr int kMaxItems = (\d+);", src).group(1)))


def cmac_item_lists(plan, G, max_heads, max_items):
    """The (RIR, partition) item lists of k_cmac, run by run, built exactly as cmac_mover_cta builds them: one window of at
    most `max_heads` RIRs at a time, partitions k_lo..k_hi of every RIR that reach the run, passes of at most `max_items`
    items. Yields (b0, nb, [(l, k, xrow_rel, mask), ...]) per pass; xrow_rel is the source block of output 0 (may be
    negative), bit s of mask says that output s takes the item."""
    K, B_valid = plan["K"], plan["B_valid"]
    irs = plan["irs"]
    for b0 in range(0, B_valid, G):
        nb = min(G, B_valid - b0)
        lmin, lmax = int(plan["lrange"][b0][0]), int(plan["lrange"][b0 + nb - 1][1])
        for l0 in range(lmin, lmax + 1, max_heads):
            heads = []
            for l in range(l0, min(l0 + max_heads, lmax + 1)):
                xb0, xnb = int(irs[l][0]), int(irs[l][1])
                d0 = b0 - xb0
                k_lo = max(0, d0 - xnb + 1)
                cnt = max(0, min(K - 1, d0 + nb - 1) - k_lo + 1) if xnb > 0 else 0
                heads.append((l, d0, xnb, k_lo, cnt))
            total = sum(h[4] for h in heads)
            flat = []
            for l, d0, xnb, k_lo, cnt in heads:
                for k in range(k_lo, k_lo + cnt):
                    jb = d0 - k
                    s_lo, s_hi = max(0, -jb), min(nb, xnb - jb)
                    mask = ((1 << s_hi) - 1) & ~((1 << s_lo) - 1)
                    flat.append((l, k, jb, mask))
            assert len(flat) == total
            for p0 in range(0, total, max_items):
                yield b0, nb, flat[p0:p0 + max_items]


def model_convolve(audio, irs, plan, scales, moving, n_out, cmac="blocks"):
    """audio (Lx,), irs (C, N, Lh), scales (N,) -> (C, n_out) following the plan. cmac="items" accumulates the output
    spectra the way k_cmac does (item lists, validity masks) instead of block by block."""
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
    if cmac == "items":
        kc = kernel_constants()
        Yall = np.zeros((B_valid, C, P + 1), dtype=complex)
        n_items = n_lists = 0
        for b0, nb, items in cmac_item_lists(plan, kc["G"], kc["max_heads"], kc["max_items"]):
            n_lists += 1
            n_items += len(items)
            for l, k, jb, mask in items:
                assert mask != 0 and mask < (1 << nb)
                for sidx in range(nb):
                    if mask >> sidx & 1:
                        assert 0 <= jb + sidx < int(plan["irs"][l][1])
                        Yall[b0 + sidx] += X[(l, jb + sidx)][None, :] * H[l, k]
        for b in range(B_valid):
            out[:, b * P:(b + 2) * P] += np.fft.irfft(Yall[b], 2 * P, axis=-1)
        res = np.zeros((C, n_out))
        res[:, :n_valid] = out[:, :n_valid]
        model_convolve.last_lists, model_convolve.last_items = n_lists, n_items
        return res
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
