"""NumPy mirror of the device synthetic generator (bigsnpr_b200/csrc/bsg_core.cu: k_synth), bit for bit.

Returns the genotype matrix (n, m) with values 0/1/2 and 3 = missing.  Test infrastructure only.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth_matrix(n, m, seed=20250924, na_rate=0.0, col_offset=0):
    j = np.arange(m, dtype=np.uint64) + np.uint64(col_offset)
    kj = mix64(np.uint64(seed) ^ mix64(j))
    maf = 0.02 + 0.48 * ((kj >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))
    thr = (maf * 16777216.0).astype(np.uint32)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        hs = mix64(kj[None, :] + i[:, None] * np.uint64(0xD1342543DE82EF95))
    a1 = (hs & np.uint64(0xFFFFFF)).astype(np.uint32) < thr[None, :]
    a2 = ((hs >> np.uint64(24)) & np.uint64(0xFFFFFF)).astype(np.uint32) < thr[None, :]
    g = a1.astype(np.uint8) + a2.astype(np.uint8)
    na_thr = np.uint32(int(na_rate * 65536.0))
    na = ((hs >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.uint32) < na_thr
    g[na] = 3
    return g


def synth_matrix_ld(n, m, seed=20250924, na_rate=0.0, col_offset=0, rho=0.9, ld_block=50):
    """NumPy mirror of k_synth_ld: the allele uniforms of a haplotype are copied from the previous SNP with
    probability rho inside blocks of ld_block global columns."""
    g = np.zeros((n, m), dtype=np.uint8)
    i = np.arange(n, dtype=np.uint64)
    na_thr = np.uint32(int(na_rate * 65536.0))
    rho_thr = np.uint32(int(rho * 65536.0))
    u0 = np.zeros(n, dtype=np.uint32)
    u1 = np.zeros(n, dtype=np.uint32)
    start = (col_offset // ld_block) * ld_block
    for gj in range(start, col_offset + m):
        kj = mix64(np.uint64(seed) ^ mix64(np.uint64(gj)))
        maf = 0.02 + 0.48 * (float(kj >> np.uint64(11)) * (1.0 / 9007199254740992.0))
        thr = np.uint32(int(maf * 16777216.0))
        with np.errstate(over="ignore"):
            hs = mix64(kj + i * np.uint64(0xD1342543DE82EF95))
        h2 = mix64(hs ^ np.uint64(0xA5A5A5A5A5A5A5A5))
        first = gj % ld_block == 0
        c0 = ((h2 & np.uint64(0xFFFF)).astype(np.uint32) < rho_thr) & (not first)
        c1 = (((h2 >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint32) < rho_thr) & (not first)
        u0 = np.where(c0, u0, (hs & np.uint64(0xFFFFFF)).astype(np.uint32))
        u1 = np.where(c1, u1, ((hs >> np.uint64(24)) & np.uint64(0xFFFFFF)).astype(np.uint32))
        if gj >= col_offset:
            col = (u0 < thr).astype(np.uint8) + (u1 < thr).astype(np.uint8)
            na = ((hs >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.uint32) < na_thr
            col[na] = 3
            g[:, gj - col_offset] = col
    return g
