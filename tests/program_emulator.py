"""numpy emulator of the lowered sweep program (test infrastructure).

Executes exactly what ``dm_sweep_kernel`` does -- tile gather, register passes with the same op
kinds and index conventions, scatter -- so the lowering stage (host C++) and the Pauli-transfer
algebra can be checked against the oracle on a CPU-only box.  The CUDA kernels themselves are
checked by the ``-m gpu`` tests.
"""
import numpy as np

CX_SRC = [0, 5, 6, 3, 4, 1, 2, 7, 11, 14, 13, 8, 15, 10, 9, 12]
CX_SGN = [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, 1, -1, 1, 1]
(K_DENSE1_A, K_DENSE1_B, K_CX_AB, K_CX_BA, K_RELAX2, K_RELAX2_SW, K_DENSE2, K_DENSE2_SW,
 K_AFF1_A, K_AFF1_B, K_ROTZ_A, K_ROTZ_B) = range(12)


def _swap_view(v):  # v[da + 4 db] -> index with (q0,q1) = (b,a)
    return v.reshape(4, 4, -1).transpose(1, 0, 2).reshape(16, -1)


def apply_op(v, kind, m):
    """v: (16, G) array, row index da + 4*db."""
    if kind == K_DENSE1_A:
        a = m[:16].reshape(4, 4)
        return np.einsum("ij,bjg->big", a, v.reshape(4, 4, -1)).reshape(16, -1)
    if kind == K_DENSE1_B:
        a = m[:16].reshape(4, 4)
        return np.einsum("ij,jag->iag", a, v.reshape(4, 4, -1)).reshape(16, -1)
    if kind in (K_AFF1_A, K_AFF1_B):
        a = np.vstack([[1.0, 0.0, 0.0, 0.0], m[:12].reshape(3, 4)])
        return apply_op(v, K_DENSE1_A if kind == K_AFF1_A else K_DENSE1_B, a.reshape(-1))
    if kind in (K_ROTZ_A, K_ROTZ_B):
        c, s = m[0], m[1]
        a = np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])
        return apply_op(v, K_DENSE1_A if kind == K_ROTZ_A else K_DENSE1_B, a.reshape(-1))
    if kind in (K_CX_AB, K_CX_BA):
        w = v if kind == K_CX_AB else _swap_view(v)
        out = np.empty_like(w)
        for i in range(16):
            out[i] = CX_SGN[i] * w[CX_SRC[i]]
        return out if kind == K_CX_AB else _swap_view(out)
    if kind in (K_RELAX2, K_RELAX2_SW):
        w = v if kind == K_RELAX2 else _swap_view(v)
        out = m[:16, None] * w
        for b in range(4):
            out[3 + 4 * b] += m[16 + b] * w[4 * b]
        for a in range(4):
            out[a + 12] += m[20 + a] * w[a]
        out[15] += m[24] * w[0]
        return out if kind == K_RELAX2 else _swap_view(out)
    if kind in (K_DENSE2, K_DENSE2_SW):
        w = v if kind == K_DENSE2 else _swap_view(v)
        out = m[:256].reshape(16, 16) @ w
        return out if kind == K_DENSE2 else _swap_view(out)
    raise ValueError(kind)


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
    mats = prog["mats"]
    op_begin = 0
    pass_begin = 0
    t = state.reshape((4,) * nd)  # axis k <-> digit nd-1-k
    for sw in prog["sweeps"]:
        assert sw[0] == pass_begin
        pos = list(sw[1:9])
        pass_end = sw[9]
        # number of tile slots = positions strictly ascending prefix
        kq = 1
        while kq < 8 and kq < nd and pos[kq] > pos[kq - 1]:
            kq += 1
        pos = pos[:kq]
        for p in range(pass_begin, pass_end):
            sa, sb, op_end = prog["passes"][p]
            da, db = pos[sa], pos[sb]
            # gather: rows da + 4 db
            axes = [nd - 1 - db, nd - 1 - da]
            tt = np.moveaxis(t, axes, [0, 1])
            shp = tt.shape
            v = tt.reshape(16, -1)  # index db*4 + da  == da + 4 db
            for o in range(op_begin, op_end):
                kind = int(prog["ops"][o][0]) & 0xff
                off = int(prog["ops"][o][1])
                v = apply_op(v, kind, mats[off:off + 256])
            t = np.moveaxis(v.reshape(shp), [0, 1], axes)
            op_begin = op_end
        pass_begin = pass_end
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
