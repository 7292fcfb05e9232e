"""numpy emulator of the lowered sweep program (test infrastructure).

Executes exactly what ``dm_sweep_kernel`` does -- tile gather, register passes with the same op
kinds and index conventions, scatter -- so the lowering stage (host C++) and the Pauli-transfer
algebra can be checked against the oracle on a CPU-only box.  The CUDA kernels themselves are
checked by the ``-m gpu`` tests.
"""
import numpy as np

CX_SRC = [0, 5, 6, 3, 4, 1, 2, 7, 11, 14, 13, 8, 15, 10, 9, 12]
CX_SGN = [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, 1, -1, 1, 1]
P_NONE, P_ROT, P_AFF, P_DENSE = range(4)
Q_NONE, Q_CXN_AB, Q_CXN_BA, Q_CX_AB, Q_CX_BA, Q_RELAX, Q_RELAX_SW, Q_DENSE, Q_DENSE_SW = range(9)


def _swap_view(v):  # v[da + 4 db] -> index with (q0,q1) = (b,a)
    return v.reshape(4, 4, -1).transpose(1, 0, 2).reshape(16, -1)


def _one_qubit(v, a, on_b):
    t = v.reshape(4, 4, -1)  # [db, da, g]
    if on_b:
        return np.einsum("ij,jag->iag", a, t).reshape(16, -1)
    return np.einsum("ij,bjg->big", a, t).reshape(16, -1)


def apply_pre(v, kind, m, on_b):
    if kind == P_NONE:
        return v
    if kind == P_ROT:  # three shears + sign, exactly as the kernel evaluates it
        t, s, sign = m[0], m[1], m[2]
        sh_x = np.array([[1, 0, 0, 0], [0, 1, -t, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        sh_y = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, s, 1, 0], [0, 0, 0, 1.0]])
        a = sh_x @ sh_y @ sh_x
        if sign < 0:
            a[1:3] *= -1.0
        return _one_qubit(v, a, on_b)
    if kind == P_AFF:
        return _one_qubit(v, np.vstack([[1.0, 0.0, 0.0, 0.0], m[:12].reshape(3, 4)]), on_b)
    if kind == P_DENSE:
        return _one_qubit(v, m[:16].reshape(4, 4), on_b)
    raise ValueError(kind)


def _cx(w):
    out = np.empty_like(w)
    for i in range(16):
        out[i] = CX_SGN[i] * w[CX_SRC[i]]
    return out


def _relax(w, m):
    out = m[:16, None] * w
    for b in range(4):
        out[3 + 4 * b] += m[16 + b] * w[4 * b]
    for a in range(4):
        out[a + 12] += m[20 + a] * w[a]
    out[15] += m[24] * w[0]
    return out


def apply_two(v, kind, m):
    """v: (16, G) array, row index da + 4*db."""
    if kind == Q_NONE:
        return v
    sw = kind in (Q_CXN_BA, Q_CX_BA, Q_RELAX_SW, Q_DENSE_SW)
    w = _swap_view(v) if sw else v
    if kind in (Q_CXN_AB, Q_CXN_BA):
        out = _relax(_cx(w), m)
    elif kind in (Q_CX_AB, Q_CX_BA):
        out = _cx(w)
    elif kind in (Q_RELAX, Q_RELAX_SW):
        out = _relax(w, m)
    elif kind in (Q_DENSE, Q_DENSE_SW):
        out = m[:256].reshape(16, 16) @ w
    else:
        raise ValueError(kind)
    return _swap_view(out) if sw else out


def decode_block(prog, sw):
    """Decodes one sweep block (layout: ml_qem_b200/csrc/program.h) -> list of passes
    [(sa, sb, [(pre_a, pre_b, twoq, off_a, off_b, off_2), ...])] and the block as doubles."""
    words = prog["prog"]
    blk = words[2 * int(sw[0]): 2 * (int(sw[0]) + int(sw[9]))]
    b = blk.view(np.uint8)
    n_passes = int(blk[:1].view(np.int32)[0])
    passes = []
    ext_q16 = int(blk[:1].view(np.int32)[1])
    for p in range(n_passes):
        h = b[16 * (1 + p): 16 * (2 + p)]
        ops_q16, n_ops = int(h[:2].view(np.uint16)[0]), int(h[2:4].view(np.uint16)[0])
        sa, sb = int(h[4]), int(h[5])
        if ext_q16:
            check_tma_pass(h, b[16 * ext_q16 + 64 * p: 16 * ext_q16 + 64 * (p + 1)].view(np.uint32), sa, sb)
        ops = []
        for o in range(n_ops):
            r = b[16 * (ops_q16 + o): 16 * (ops_q16 + o + 1)]
            offs = r[4:10].view(np.uint16)
            ops.append((int(r[0]), int(r[1]), int(r[2]), int(offs[0]), int(offs[1]), int(offs[2])))
        passes.append((sa, sb, ops))
    return passes, blk.view(np.float64)


def tswz(j):
    """CU_TENSOR_MAP_SWIZZLE_128B on 8-byte element indices (program.h)."""
    return j ^ (((j >> 4) & 7) << 1)


def tma_pass_layout(sa, sb):
    """numpy restatement of program.h tma_pass_layout: thread-id bit -> tile index bit, beta bit."""
    is_target = [(b >> 1) in (sa, sb) for b in range(12)]
    used = list(is_target)
    slot0 = sa == 0 or sb == 0
    beta = 0
    if not slot0:
        used[0] = True
    tbit = [-1] * 7
    for k in (1, 2, 3):
        pick = k if not used[k] else (k + 3 if not used[k + 3] else -1)
        tbit[k - 1] = pick
        if pick >= 0:
            used[pick] = True
    if slot0:
        beta = max(b for b in range(12) if not used[b])
        used[beta] = True
    free = [b for b in range(12) if not used[b]]
    for k in range(7):
        if tbit[k] < 0:
            tbit[k] = free.pop(0)
    return tbit, beta


def check_tma_pass(h, corners, sa, sb):
    """The host-computed addressing of a TMA-layout pass must (1) cover every tile element exactly
    once over 128 threads x 2 groups x 16 corners and (2) keep the 16-byte accesses of the fast
    layout aligned; returns the worst quarter-warp bank-conflict degree of the gathers."""
    row, gofs = int(h[8]), int(h[12:16].view(np.uint32)[0])
    lo, hi = min(sa, sb), max(sa, sb)
    assert row == hi * (hi - 1) // 2 + lo
    tbit, beta = tma_pass_layout(lo, hi)
    assert gofs == 8 * tswz(1 << beta)
    tid = np.arange(128)
    j = np.zeros(128, dtype=np.int64)
    for k in range(7):
        j |= ((tid >> k) & 1) << tbit[k]
    base = 8 * tswz(j)
    assert [int(c) for c in corners] == [8 * tswz((i & 3) << (2 * sa) | (i >> 2) << (2 * sb)) for i in range(16)]
    seen = np.zeros(4096, dtype=np.int32)
    worst = 1
    for c in corners:
        a0 = base ^ int(c)
        for g in (0, 1):
            np.add.at(seen, (a0 ^ (gofs if g else 0)) // 8, 1)
        if gofs == 8:
            assert not (a0 & 15).any()
            for q in range(0, 128, 8):  # quarter warp: 8 lanes x 16 B must hit 8 distinct chunks
                worst = max(worst, 8 // len(set(((a0[q:q + 8] >> 4) & 7).tolist())))
    assert (seen == 1).all()
    return worst


def run_program(prog):
    """Returns the final Pauli-basis state r (4^n doubles) of a lowered program (engine.lower_dm)."""
    nd = prog["n_digits"]
    state = np.zeros(4 ** nd)
    idx = np.arange(4 ** nd)
    ok = np.ones(4 ** nd, dtype=bool)
    for d in range(nd):
        dig = (idx >> (2 * d)) & 3
        ok &= (dig == 0) | (dig == 3)
    state[ok] = 1.0
    t = state.reshape((4,) * nd)  # axis k <-> digit nd-1-k
    for sw in prog["sweeps"]:
        pos = list(sw[1:9])
        if pos[7] == 0x40:  # TMA tile layout (program.h): slots 2..5 follow the box order in pos[6]
            pos = pos[:2] + [pos[2 + ((pos[6] >> (2 * k)) & 3)] for k in range(4)]
        else:
            kq = 1
            while kq < 7 and kq < nd and pos[kq] > pos[kq - 1]:  # pos[7] is the first-pass descriptor
                kq += 1
            pos = pos[:kq]
        passes, mats = decode_block(prog, sw)
        for sa, sb, ops in passes:
            da, db = pos[sa], pos[sb]
            axes = [nd - 1 - db, nd - 1 - da]
            tt = np.moveaxis(t, axes, [0, 1])
            shp = tt.shape
            v = tt.reshape(16, -1)  # index db*4 + da == da + 4 db
            for pre_a, pre_b, twoq, off_a, off_b, off_2 in ops:
                v = apply_pre(v, pre_a, mats[off_a:off_a + 16], False)
                v = apply_pre(v, pre_b, mats[off_b:off_b + 16], True)
                v = apply_two(v, twoq, mats[off_2:off_2 + 256])
            t = np.moveaxis(v.reshape(shp), [0, 1], axes)
    return np.ascontiguousarray(t).reshape(-1)


def expvals(prog, state, obs_term_counts):
    vals = []
    k = 0
    for cnt in obs_term_counts:
        tot = 0.0
        for _ in range(cnt):
            i = prog["term_index"][k]
            if i >= 0:
                tot += prog["term_coeff"][k] * state[i]
            k += 1
        vals.append(tot)
    return np.array(vals)
