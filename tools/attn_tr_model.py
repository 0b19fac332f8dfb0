"""Lane-level model of the transposed attention tile (paged_attn_persist_kernel<…, TR = 1>).

The default stream kernel computes S = Q K^T with the (padded) 16 query rows on the MMA's M.  For
decode G * q_len <= 8 rows, so the transposed product wastes less and needs half the registers:

    S^T [16 keys x 8 rows]  = K  [16 keys x D]  . Q^T [D x 8 rows]      (A = K via ldmatrix, B = Q regs)
    O^T [D x 8 rows]       += V^T [D x 16 keys] . P^T [16 keys x 8 rows] (A = V via ldmatrix.trans,
                                                                        B = P^T via movmatrix.trans)

This file restates `mma.m16n8k16`, `ldmatrix(.trans)` and `movmatrix.trans` at the level of the 32
lanes' registers and runs the kernel's per-tile arithmetic with exactly the fragment indexing the
CUDA code uses, against a plain softmax(QK^T)V.  It guards the index algebra of the kernel
(which lane holds which element, which shuffles reduce a column), not the compiled code.

  python tools/attn_tr_model.py
"""
import numpy as np

LANES = 32


def bf16(x):
    """round-to-nearest-even to bfloat16, kept in float32"""
    a = np.asarray(x, dtype=np.float32).copy()
    u = a.view(np.uint32)
    u += 0x7FFF + ((u >> 16) & 1)
    u &= 0xFFFF0000
    return u.view(np.float32)


# ---- warp-level primitives -------------------------------------------------------------------
def mma_16816(d, a, b):
    """d[lane][4] += A[16x16] . B[16x8]; a[lane][4][2], b[lane][2][2] hold bf16 pairs."""
    A = np.zeros((16, 16), np.float32)
    B = np.zeros((16, 8), np.float32)
    for L in range(LANES):
        g, t = L >> 2, L & 3
        for e in range(2):
            A[g, 2 * t + e] = a[L][0][e]
            A[g + 8, 2 * t + e] = a[L][1][e]
            A[g, 2 * t + 8 + e] = a[L][2][e]
            A[g + 8, 2 * t + 8 + e] = a[L][3][e]
            B[2 * t + e, g] = b[L][0][e]
            B[2 * t + 8 + e, g] = b[L][1][e]
    C = A.astype(np.float64) @ B.astype(np.float64)
    for L in range(LANES):
        g, t = L >> 2, L & 3
        d[L][0] += C[g, 2 * t]
        d[L][1] += C[g, 2 * t + 1]
        d[L][2] += C[g + 8, 2 * t]
        d[L][3] += C[g + 8, 2 * t + 1]


def ldmatrix_x4(mem, row_of_lane, trans):
    """mem[row][col] (a 2-D view of shared memory in elements); row_of_lane[L] = (row, col0) of the
    8-element row that lane L addresses: lanes 8j..8j+7 give the rows of matrix j.  Returns
    r[lane][4][2]."""
    mats = []
    for j in range(4):
        m = np.stack([mem[row_of_lane[8 * j + i][0], row_of_lane[8 * j + i][1]:row_of_lane[8 * j + i][1] + 8]
                      for i in range(8)])
        mats.append(m.T if trans else m)
    return [[[mats[j][L >> 2, 2 * (L & 3) + e] for e in range(2)] for j in range(4)] for L in range(LANES)]


def movmatrix_trans(r):
    """r[lane][2]: lane L holds M[L >> 2][2 (L & 3) + e]; returns the fragment of M^T."""
    M = np.zeros((8, 8), np.float32)
    for L in range(LANES):
        for e in range(2):
            M[L >> 2, 2 * (L & 3) + e] = r[L][e]
    return [[M[2 * (L & 3) + e, L >> 2] for e in range(2)] for L in range(LANES)]


def shfl_xor(vals, mask):
    return [vals[L ^ mask] for L in range(LANES)]


# ---- the transposed tile loop, as the kernel runs it ---------------------------------------------
def attend_transposed(q, k, v, kv_end, scale_log2, D):
    """q: [n_rows <= 8][D] bf16 values; k, v: [n_tiles * 16][D]; keys >= kv_end are masked (the
    causal end of a decode row).  Returns (O [n_rows][D] fp32 normalised, lse2 [n_rows])."""
    n_rows = q.shape[0]
    KS, MB = D // 16, D // 16
    n_tiles = k.shape[0] // 16
    # B fragments of Q^T: lane (g, t) holds Q[row g][ks*16 + 2t, +1] and [.. + 8, + 9]
    qb = [[[[q[L >> 2, ks * 16 + 2 * (L & 3) + 8 * h + e] if (L >> 2) < n_rows else 0.0
             for e in range(2)] for h in range(2)] for ks in range(KS)] for L in range(LANES)]
    m = [[-np.inf, -np.inf] for _ in range(LANES)]      # per column c: query row 2t + c
    l = [[0.0, 0.0] for _ in range(LANES)]
    ot = [[[0.0] * 4 for _ in range(MB)] for _ in range(LANES)]
    for ti in range(n_tiles):
        kt, vt = k[ti * 16:(ti + 1) * 16], v[ti * 16:(ti + 1) * 16].copy()
        pos0 = ti * 16
        if pos0 + 16 > kv_end:
            vt[max(0, kv_end - pos0):] = 0.0             # the kernel zeroes V rows past the end
        sacc = [[0.0] * 4 for _ in range(LANES)]
        for ks in range(KS):
            # A = K[16 keys][16 dims]: matrices (keys 0-7 | 8-15) x (dims 0-7 | 8-15) -> a0..a3
            rows = [(((L >> 3) & 1) * 8 + (L & 7), ks * 16 + (L >> 4) * 8) for L in range(LANES)]
            a = ldmatrix_x4(kt, rows, trans=False)
            mma_16816(sacc, a, [qb[L][ks] for L in range(LANES)])
        # online softmax down the key axis: lane (g, t) holds keys {g, g + 8} x rows {2t, 2t + 1}
        x = [[0.0] * 4 for _ in range(LANES)]
        mx = [[m[L][0], m[L][1]] for L in range(LANES)]
        for L in range(LANES):
            g = L >> 2
            for c in range(2):
                for hh in range(2):                       # element index 2 * hh + c: key g + 8 hh
                    pos = pos0 + g + 8 * hh
                    val = sacc[L][2 * hh + c] * scale_log2 if pos < kv_end else -np.inf
                    x[L][2 * hh + c] = val
                    mx[L][c] = max(mx[L][c], val)
        for mask in (4, 8, 16):                           # reduce over g (lane bits 2..4)
            for c in range(2):
                col = [mx[L][c] for L in range(LANES)]
                sh = shfl_xor(col, mask)
                for L in range(LANES):
                    mx[L][c] = max(mx[L][c], sh[L])
        corr = [[1.0, 1.0] for _ in range(LANES)]
        for L in range(LANES):
            for c in range(2):
                ms = 0.0 if mx[L][c] == -np.inf else mx[L][c]
                corr[L][c] = 1.0 if mx[L][c] == m[L][c] else float(np.exp2(np.float32(m[L][c] - ms)))
                s = 0.0
                for hh in range(2):
                    x[L][2 * hh + c] = float(np.exp2(np.float32(x[L][2 * hh + c] - ms)))
                    s += x[L][2 * hh + c]
                l[L][c] = l[L][c] * corr[L][c] + s
                m[L][c] = mx[L][c]
            for mb in range(MB):
                ot[L][mb][0] *= corr[L][0]
                ot[L][mb][1] *= corr[L][1]
                ot[L][mb][2] *= corr[L][0]
                ot[L][mb][3] *= corr[L][1]
        # P^T -> B fragment: pack (key g | g + 8; rows 2t, 2t + 1), transpose each 8x8 block
        lo = movmatrix_trans([[bf16(x[L][0]), bf16(x[L][1])] for L in range(LANES)])
        hi = movmatrix_trans([[bf16(x[L][2]), bf16(x[L][3])] for L in range(LANES)])
        pb = [[lo[L], hi[L]] for L in range(LANES)]       # b0: keys 2t, 2t+1 of row g; b1: keys + 8
        for mb in range(MB):
            # A = V^T[16 dims][16 keys] via ldmatrix.trans of (keys 0-7 | 8-15) x (dims 0-7 | 8-15):
            # a0 (dims lo, keys lo), a1 (dims hi, keys lo), a2 (dims lo, keys hi), a3 (dims hi, keys hi)
            rows = [((L >> 4) * 8 + (L & 7), mb * 16 + ((L >> 3) & 1) * 8) for L in range(LANES)]
            a = ldmatrix_x4(vt, rows, trans=True)
            acc = [ot[L][mb] for L in range(LANES)]
            mma_16816(acc, a, pb)
    # finalize: l over g, then lane (g, t) owns O^T[dim mb*16 + g (+8)][row 2t + c]
    for mask in (4, 8, 16):
        for c in range(2):
            col = [l[L][c] for L in range(LANES)]
            sh = shfl_xor(col, mask)
            for L in range(LANES):
                l[L][c] += sh[L]
    O = np.zeros((n_rows, D), np.float32)
    lse = np.zeros(n_rows, np.float32)
    for L in range(LANES):
        g, t = L >> 2, L & 3
        for c in range(2):
            r = 2 * t + c
            if r >= n_rows:
                continue
            for mb in range(MB):
                O[r, mb * 16 + g] = ot[L][mb][c] / l[L][c]
                O[r, mb * 16 + g + 8] = ot[L][mb][2 + c] / l[L][c]
            lse[r] = m[L][c] + np.log2(l[L][c])
    return O, lse


def reference(q, k, v, kv_end, scale_log2):
    s = (q.astype(np.float64) @ k[:kv_end].astype(np.float64).T) * scale_log2
    mx = s.max(axis=1, keepdims=True)
    p = np.exp2(s - mx)
    return (p @ v[:kv_end].astype(np.float64)) / p.sum(axis=1, keepdims=True), (mx[:, 0] + np.log2(p.sum(axis=1)))


def check(seed=0, D=128, n_rows=4, n_tiles=5, kv_end=None):
    rng = np.random.default_rng(seed)
    q = bf16(rng.standard_normal((n_rows, D)))
    k = bf16(rng.standard_normal((n_tiles * 16, D)))
    v = bf16(rng.standard_normal((n_tiles * 16, D)))
    if kv_end is not None:                       # whatever lies past the end must not matter
        v[kv_end:] = np.nan
        k[kv_end:] = 1e30
    kv_end = n_tiles * 16 if kv_end is None else kv_end
    sc = D ** -0.5 * 1.4426950408889634
    got, lse = attend_transposed(q, k, np.nan_to_num(v, nan=0.0) if False else v, kv_end, sc, D)
    want, lse_w = reference(q, k, v, kv_end, sc)
    err = np.abs(got - want).max()
    assert np.isfinite(got).all() and err < 2e-2, err   # P is rounded to bf16 before PV, like the kernel
    assert np.abs(lse - lse_w).max() < 1e-3
    return err


if __name__ == "__main__":
    for n_rows in (1, 4, 7, 8):
        print("rows", n_rows, "max |err|", check(seed=n_rows, n_rows=n_rows))
    print("ragged end", check(seed=9, n_rows=4, n_tiles=4, kv_end=53))
    print("D=64", check(seed=3, D=64, n_rows=2, n_tiles=3, kv_end=40))
