"""CPU oracle for the AMB path -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py for the rules).

numpy restatement of the reference's CSR -> AMB conversion and of its SpMV kernel's decode, step
by step, so the GPU conversion can be compared ARRAY BY ARRAY (bit-exact: it is integer / copy
work) and the GPU SpMV value by value.  Paths are relative to the nsparse tree.

    convert_amb()   cuda-c/src/conversion/convert_amb.cu:604-833 (convert_amb_at) for one fixed
                    (seg_size, block_size):
                      segmented CSR            :138-251
                      sigma-window stable sort :660-696
                      chunk widths / offsets   :46-102
                      column-major ELL fill    :104-136
                      16-bit columns, cl pack  :313-346, empty-chunk packing :348-386
                      write permutation        :253-299
                      column blocking          :388-525
    spmv_decode()   cuda-c/src/kernel/kernel_spmv_amb.cu:21-79 (kernel_spmv_amb_atomic)
    plan_footprint() the reference's own footprint model (convert_amb.cu:785-797, the non-`AT`
                    build) over its candidate segment sizes (:879-892) and block sizes 1..20.

Dense in (pad_M x seg_num) like the reference: only for test-sized inputs.

Deliberate deviations (both are defects of the reference):
  * the sigma windows are ALWAYS sorted.  The reference skips a window when check_nnz[...] == 0 but
    indexes that array with div_round_up(pad_M, min(SIGMA, M)) instead of
    div_round_up(pad_M, SIGMA) (:686 vs :553), which reads a wrong / out-of-bounds flag whenever
    M < 32768 and M % 32 != 0.  Sorting an all-empty window is the identity, so always sorting is
    what the reference computes whenever its flag is read correctly.
  * unsorted / duplicate columns.  The reference assumes strictly ascending columns inside a row:
    set_blocked_cl (:406-411) treats a column that is not larger than the block base as "same block"
    while set_blocked_col_val (:500-509) cannot place it, so entries are silently LOST (for every
    block size, including 1).  Here the entries of a virtual row are first put in ascending column
    order, and a repeated column opens a new block in the count exactly where the fill needs one.
    On ascending distinct columns both changes are the identity, so the arrays are the
    reference's.
  * spmv_decode() never reads x beyond N (the reference reads x[c + b] up to N + 19 and relies on the
    caller's padding, kernel_spmv_amb.cu:51-62; the stored value there is always 0).

Parity status: pinned against the reference's own convert_amb.cu / kernel_spmv_amb.cu (the
`__shfl_xor` -> `__shfl_xor_sync` patched copy built by oracle/ref_gpu/Makefile, run on a B200) through
the fixtures in tests/golden/amb_*.npz, see tests/golden/make_amb_golden.py.
"""
from __future__ import annotations

import numpy as np

CHUNK = 32            # WARP, sf_csr2amb: mat->chunk = WARP (convert_amb.cu:859)
SIGMA = 32768         # SHORT_MAX (:863)
USHORT_MAX = 65536
SCL_BORDER = 16
SCL_BIT = 0xFFFF
MAX_BLOCK_SIZE = 20


def _div_up(a, b):
    return (a + b - 1) // b


def convert_amb(rpt, col, val, M, N, seg_size, block_size, sigma=SIGMA):
    rpt = np.asarray(rpt, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val)
    seg_size = int(seg_size)
    bs = int(block_size)
    pad_M = CHUNK * _div_up(M, CHUNK)
    G = _div_up(N, seg_size)
    total = pad_M * G

    # ---- segmented CSR: virtual row v = g * pad_M + i, entries keep their CSR order (:138-206) ----
    row_of = np.repeat(np.arange(M, dtype=np.int64), np.diff(rpt))
    key = (col // seg_size) * pad_M + row_of
    # (deviation, see the module docstring) entries of a virtual row are put in ascending column order
    order = np.lexsort((col, key))
    seg_col, seg_val = col[order], val[order]
    nnz_num = np.bincount(key, minlength=total).astype(np.int64)
    seg_rpt = np.concatenate([[0], np.cumsum(nnz_num)])

    # ---- stable sort by nnz (descending) inside every window of min(sigma, M) rows (:660-696) ----
    perm = np.arange(total, dtype=np.int64)
    S = min(sigma, M)
    if S > 1:
        for g in range(G):
            for start in range(0, M, S):
                end = min(start + S, M)
                lo, hi = g * pad_M + start, g * pad_M + end
                if not nnz_num[lo:hi].any():
                    continue
                o = np.argsort(-nnz_num[lo:hi], kind="stable")
                nnz_num[lo:hi] = nnz_num[lo:hi][o]
                perm[lo:hi] = perm[lo:hi][o]

    # ---- chunk width = longest row of the chunk, offsets = scan (:46-102) ----
    full_cl = nnz_num.reshape(-1, CHUNK).max(axis=1)
    full_cs = np.concatenate([[0], np.cumsum(full_cl * CHUNK)])[:-1]
    nnz_ell = int((full_cl * CHUNK).sum())

    # ---- column-major ELL fill; padding takes the column of the chunk's FIRST row, value 0 (:104-136) ----
    p = np.arange(nnz_ell, dtype=np.int64)
    b = np.searchsorted(full_cs + full_cl * CHUNK, p, side="right")      # chunk of slot p
    r = p - full_cs[b]
    tid, j = r % CHUNK, r // CHUNK
    i = b * CHUNK + tid
    real = j < nnz_num[i]
    src = np.where(real, seg_rpt[perm[i]] + j, seg_rpt[perm[b * CHUNK]] + j)
    ell_col = seg_col[src]
    ell_val = np.where(real, seg_val[src], 0).astype(val.dtype)

    # ---- 16-bit columns, segment id into cl, empty chunks dropped (:301-386) ----
    us_col = (ell_col % seg_size).astype(np.int64)
    nonempty = full_cl != 0
    c_size = int(nonempty.sum())
    first_col = np.zeros(len(full_cl), dtype=np.int64)
    first_col[nonempty] = ell_col[full_cs[nonempty]]
    packed_cl = ((full_cl - 1) | ((first_col // seg_size) << SCL_BORDER))[nonempty]
    packed_cs = full_cs[nonempty]

    # ---- write permutation: row id of every lane, 16-bit low part + per-chunk high part (:253-299) ----
    wp = perm - (np.arange(total, dtype=np.int64) // pad_M) * pad_M
    write_perm = wp.reshape(-1, CHUNK)[nonempty].reshape(-1)
    s_write_perm = (write_perm % USHORT_MAX).astype(np.uint16)
    s_write_off = (write_perm.reshape(-1, CHUNK)[:, 0] // USHORT_MAX).astype(np.uint16)

    # ---- per-lane view of the unblocked ELL rows: [c_size * 32, maxw] ----
    clw = (packed_cl & SCL_BIT) + 1
    maxw = int(clw.max()) if c_size else 0
    lanes = c_size * CHUNK
    lane_chunk = np.repeat(np.arange(c_size), CHUNK)
    lane_tid = np.tile(np.arange(CHUNK), c_size)
    lane_w = clw[lane_chunk]
    lane_cnt = nnz_num.reshape(-1, CHUNK)[nonempty].reshape(-1)
    kk = np.arange(maxw)
    valid = kk[None, :] < lane_w[:, None]
    addr = packed_cs[lane_chunk][:, None] + lane_tid[:, None] + kk[None, :] * CHUNK
    addr = np.where(valid, addr, 0)
    lc = np.where(valid, us_col[addr] if nnz_ell else 0, 0)
    lv = np.where(valid, ell_val[addr] if nnz_ell else 0, 0).astype(val.dtype)

    # ---- set_blocked_cl (:388-429): blocks needed by the lane, max over the chunk ----
    base = lc[:, 0].copy() if maxw else np.zeros(lanes, np.int64)
    width = np.zeros(lanes, dtype=np.int64)
    for k in range(1, maxw):
        # the second term never fires on ascending distinct columns (deviation, module docstring)
        cond = valid[:, k] & ((lc[:, k] - base >= bs) | ((k < lane_cnt) & (lc[:, k] <= lc[:, k - 1])))
        base = np.where(cond, lc[:, k], base)
        width += bs * cond
    width += bs
    blocks = width.reshape(-1, CHUNK).max(axis=1) // bs if c_size else np.zeros(0, np.int64)
    blocked_cl = ((blocks - 1) | ((packed_cl >> SCL_BORDER) << SCL_BORDER)).astype(np.uint32)
    blocked_cs = np.concatenate([[0], np.cumsum(blocks * CHUNK * bs)])[:-1].astype(np.int64)
    c_nnz = int((blocks * CHUNK * bs).sum())

    # ---- set_blocked_col_val (:473-525), the per-lane state machine vectorised over lanes ----
    b_col = np.zeros(c_nnz // bs, dtype=np.uint16)
    b_val = np.zeros(c_nnz, dtype=val.dtype)
    it = np.zeros(lanes, dtype=np.int64)
    lane_blocks = blocks[lane_chunk] if c_size else np.zeros(0, np.int64)
    cbase = blocked_cs[lane_chunk] if c_size else np.zeros(0, np.int64)
    ar = np.arange(lanes)
    maxb = int(blocks.max()) if c_size else 0
    for k in range(maxb):
        act = k < lane_blocks
        have = act & (it < lane_w)
        itc = np.minimum(it, np.maximum(lane_w - 1, 0))
        c = lc[ar, itc]
        last = lc[ar, np.maximum(lane_w - 1, 0)]
        bcol = np.where(have, c, (last // bs) * bs)
        pos = cbase // bs + lane_tid + k * CHUNK
        b_col[pos[act]] = bcol[act].astype(np.uint16)
        # h = 0 holds the block's first entry (c - base == 0)
        vpos = cbase + lane_tid + (k * bs) * CHUNK
        b_val[vpos[have]] = lv[ar, itc][have]
        it = it + have
        blk_base = c
        for h in range(1, bs):
            itc = np.minimum(it, np.maximum(lane_w - 1, 0))
            hit = have & (it < lane_w) & (lc[ar, itc] - blk_base == h)
            vpos = cbase + lane_tid + (k * bs + h) * CHUNK
            b_val[vpos[hit]] = lv[ar, itc][hit]
            it = it + hit

    return dict(
        M=M, N=N, pad_M=pad_M, chunk=CHUNK, SIGMA=sigma, seg_size=seg_size, seg_num=G, group_num_col=G,
        block_size=bs, c_size=c_size, nnz=c_nnz, nnz_unblocked=nnz_ell,
        cs=blocked_cs.astype(np.int32), cl=blocked_cl,
        sellcs_col=b_col, sellcs_val=b_val,
        s_write_permutation=s_write_perm, s_write_permutation_offset=s_write_off,
        write_permutation=write_perm.astype(np.int32),
    )


def spmv_decode(amb, x):
    """y = A x from the AMB arrays, following kernel_spmv_amb_atomic lane by lane (fp64 accumulate
    inside a lane in `real`, like the kernel; lanes are summed into y in lane order)."""
    x = np.asarray(x)
    dt = amb["sellcs_val"].dtype
    bs, c_size, seg = amb["block_size"], amb["c_size"], amb["seg_size"]
    lanes = c_size * CHUNK
    y = np.zeros(amb["pad_M"], dtype=dt)
    if lanes == 0:
        return y[:amb["M"]]
    lane_chunk = np.arange(lanes) >> 5
    tid = np.arange(lanes) & 31
    offset = amb["s_write_permutation"].astype(np.int64) + \
        amb["s_write_permutation_offset"].astype(np.int64)[lane_chunk] * USHORT_MAX
    cs = amb["cs"].astype(np.int64)[lane_chunk]
    cl = amb["cl"].astype(np.int64)[lane_chunk]
    width = (cl & SCL_BIT) + 1
    c_off = (cl >> SCL_BORDER) * seg
    ans = np.zeros(lanes, dtype=dt)
    N = amb["N"]
    for h in range(int(width.max())):
        act = h < width
        cpos = np.where(act, cs // bs + tid + h * CHUNK, 0)
        c = amb["sellcs_col"][cpos].astype(np.int64) + c_off
        for b in range(bs):
            vpos = np.where(act, cs + tid + (h * bs + b) * CHUNK, 0)
            v = amb["sellcs_val"][vpos]
            xi = np.minimum(c + b, N - 1)           # never read beyond N; v is 0 there
            ans = np.where(act, ans + v * x[xi].astype(dt), ans)
    np.add.at(y, offset, ans)
    return y[:amb["M"]]


def to_dense_entries(amb):
    """(row, col, val) of every stored NON-ZERO-POSITION slot whose value is not an explicit padding
    zero is not recoverable (real zeros look like padding); tests compare through spmv_decode and
    through per-row multisets of non-zero values instead.  Returns rows/cols/vals of all slots."""
    bs, c_size, seg = amb["block_size"], amb["c_size"], amb["seg_size"]
    lanes = c_size * CHUNK
    lane_chunk = np.arange(lanes) >> 5
    tid = np.arange(lanes) & 31
    row = amb["s_write_permutation"].astype(np.int64) + \
        amb["s_write_permutation_offset"].astype(np.int64)[lane_chunk] * USHORT_MAX
    cs = amb["cs"].astype(np.int64)[lane_chunk]
    cl = amb["cl"].astype(np.int64)[lane_chunk]
    width = (cl & SCL_BIT) + 1
    c_off = (cl >> SCL_BORDER) * seg
    R, Cc, Vv = [], [], []
    for h in range(int(width.max()) if lanes else 0):
        act = h < width
        c = amb["sellcs_col"][(cs // bs + tid + h * CHUNK)[act]].astype(np.int64) + c_off[act]
        for b in range(bs):
            v = amb["sellcs_val"][(cs + tid + (h * bs + b) * CHUNK)[act]]
            R.append(row[act])
            Cc.append(c + b)
            Vv.append(v)
    if not R:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)
    return np.concatenate(R), np.concatenate(Cc), np.concatenate(Vv)


def seg_candidates(N):
    """Candidate segment sizes of sf_csr2amb's search (convert_amb.cu:879-892)."""
    if N >= 128 * 1024:
        return [64 * 1024]
    if N < 100:
        return [64 * 1024, 1, 2, 3, 4]
    return [64 * 1024, 1024, 2048, 3072, 4096]


def footprint(amb, real_bytes):
    """convert_amb.cu:785-791 (int arithmetic)."""
    f = (amb["nnz"] // amb["block_size"]) * 2
    f += amb["nnz"] * real_bytes
    f += amb["c_size"] * 4 * 2
    f += amb["c_size"] * CHUNK * 2 + amb["c_size"] * 2
    f += amb["c_size"] * CHUNK * real_bytes * 2
    f += amb["M"] * real_bytes * 2
    return f


def plan_footprint(rpt, col, val, M, N):
    """(seg_size, block_size) minimising the footprint model; the first minimum wins (strict `>`
    at :793), segment sizes in seg_candidates order, block sizes ascending."""
    best = None
    rb = np.asarray(val).dtype.itemsize
    for seg in seg_candidates(N):
        for bs in range(1, MAX_BLOCK_SIZE + 1):
            amb = convert_amb(rpt, col, val, M, N, seg, bs)
            f = footprint(amb, rb)
            if best is None or best[0] > f:
                best = (f, seg, bs)
    return best[1], best[2], best[0]
