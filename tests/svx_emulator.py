"""numpy emulator of the statevector sweep program (test infrastructure): executes the segments of
an SvxProgram -- tile sweeps with the kernel's op kinds and index conventions, the EXCHANGE
all-to-all between simulated ranks, Z-type expectation sums -- so the planner (host C++) can be
checked against the oracle on a CPU-only box, including the amplitude-sharded case."""
import numpy as np

SVO_U1, SVO_X, SVO_U2, SVO_SWAP, SVO_D1, SVO_D2, SVO_R1, SVO_X1, SVO_DZZ = 1, 2, 3, 4, 5, 6, 7, 8, 9
SVF_ON_B, SVF_COND, SVF_COND_VAL = 1, 2, 4


def decode_block(info, sw):
    words = info["prog"]
    blk = words[2 * int(sw[0]): 2 * (int(sw[0]) + (int(sw[9]) & 0xffff))]  # high half: first-pass descriptor
    b = blk.view(np.uint8)
    n_passes = int(blk[:1].view(np.int32)[0])
    passes = []
    for p in range(n_passes):
        h = b[16 + 8 * p: 24 + 8 * p]
        ops_q8, n_ops = int(h[:2].view(np.uint16)[0]), int(h[2:4].view(np.uint16)[0])
        ops = []
        for o in range(n_ops):
            r = b[8 * (ops_q8 + o): 8 * (ops_q8 + o + 1)]
            ops.append(dict(kind=int(r[0]), flags=int(r[1]), qa=int(r[2]), qb=int(r[3]),
                            off=int(r[4:6].view(np.uint16)[0]), cond_bit=int(r[6])))
        passes.append((int(h[4]), int(h[5]), int(h[6]), ops))
    return passes, blk.view(np.float64)


def _cplx(m, off, n):
    return m[off:off + 2 * n].view(np.complex128)


def apply_sweeps(info, psi, rank, first, count):
    """Executes sweeps [first, first+count) on one shard (in place)."""
    nl, K = info["n_local"], info["tile_bits"]
    LB = max(0, K - 8)
    idx = np.arange(1 << nl, dtype=np.int64)
    gi = (rank << nl) | idx
    for sw in info["sweeps"][first:first + count]:
        slotpos = list(range(LB)) + [int(x) for x in sw[1:1 + K - LB]]
        assert all(p < nl for p in slotpos) and sorted(set(slotpos)) == slotpos
        passes, mats = decode_block(info, sw)
        for sa, sb, needs_index, ops in passes:
            pa, pb = slotpos[sa], slotpos[sb]
            assert pa != pb
            for op in ops:
                cond = np.ones(1 << nl, dtype=bool)
                if op["flags"] & SVF_COND:
                    assert needs_index
                    want = 1 if op["flags"] & SVF_COND_VAL else 0
                    cond = ((gi >> op["cond_bit"]) & 1) == want
                on_b = bool(op["flags"] & SVF_ON_B)
                k = op["kind"]
                if k == SVO_DZZ:
                    assert needs_index
                    hdr = mats[op["off"]:].view(np.uint32)
                    n_d, kz = int(hdr[0]), int(hdr[1])
                    w = np.zeros_like(gi)
                    seen_pairs = 0
                    for j in range(n_d):
                        d, M = int(hdr[2 + 2 * j]), int(hdr[3 + 2 * j])
                        y = (gi ^ (gi >> d)) & M
                        seen_pairs += bin(M).count("1")
                        for q in range(32):
                            w += (y >> q) & 1
                    assert seen_pairs == kz
                    tab = _cplx(mats, op["off"] + ((n_d + 2) & ~1), kz + 1)
                    psi *= tab[w]
                elif k in (SVO_U1, SVO_X, SVO_R1, SVO_X1):
                    pt = pb if on_b else pa
                    assert not (op["flags"] & SVF_COND) or op["cond_bit"] != pt
                    i0 = idx[((idx >> pt) & 1) == 0]
                    i0 = i0[cond[i0]]
                    i1 = i0 | (1 << pt)
                    a, b = psi[i0].copy(), psi[i1].copy()
                    if k == SVO_X:
                        psi[i0], psi[i1] = b, a
                    elif k == SVO_R1:
                        assert not (op["flags"] & SVF_COND)
                        m = mats[op["off"]:op["off"] + 4]
                        psi[i0] = m[0] * a + m[1] * b
                        psi[i1] = m[2] * a + m[3] * b
                    elif k == SVO_X1:
                        assert not (op["flags"] & SVF_COND)
                        m = mats[op["off"]:op["off"] + 4]
                        psi[i0] = m[0] * a + 1j * m[2] * b
                        psi[i1] = 1j * m[3] * a + m[1] * b
                    else:
                        u = _cplx(mats, op["off"], 4)
                        psi[i0] = u[0] * a + u[1] * b
                        psi[i1] = u[2] * a + u[3] * b
                elif k in (SVO_U2, SVO_SWAP):
                    p0, p1 = (pb, pa) if on_b else (pa, pb)  # matrix index i_p0 + 2 i_p1
                    base = idx[(((idx >> p0) & 1) == 0) & (((idx >> p1) & 1) == 0)]
                    ii = [base, base | (1 << p0), base | (1 << p1), base | (1 << p0) | (1 << p1)]
                    x = np.stack([psi[i] for i in ii])
                    if k == SVO_SWAP:
                        y = x[[0, 2, 1, 3]]
                    else:
                        y = _cplx(mats, op["off"], 16).reshape(4, 4) @ x
                    for j, i in enumerate(ii):
                        psi[i] = y[j]
                elif k == SVO_D1:
                    assert needs_index
                    ph = _cplx(mats, op["off"], 2)
                    psi *= ph[(gi >> op["qa"]) & 1]
                elif k == SVO_D2:
                    assert needs_index
                    ph = _cplx(mats, op["off"], 4)
                    psi *= ph[((gi >> op["qa"]) & 1) | (((gi >> op["qb"]) & 1) << 1)]
                else:
                    raise ValueError(k)


def zexp(info, psi, rank, first, count, vals):
    """Adds this shard's contribution of Z-terms [first, first+count) into vals."""
    nl = info["n_local"]
    gi = (rank << nl) | np.arange(1 << nl, dtype=np.int64)
    prob = np.abs(psi) ** 2
    for t in range(first, first + count):
        par = np.zeros_like(gi)
        m = int(info["zt_mask"][t])
        q = 0
        while m >> q:
            if (m >> q) & 1:
                par ^= (gi >> q) & 1
            q += 1
        vals[int(info["zt_obs"][t])] += info["zt_coeff"][t] * float(np.sum((1.0 - 2.0 * par) * prob))


def run(info, world=None):
    """All ranks simulated in this process -> (values[n_observables], final shards)."""
    nl, g = info["n_local"], info["n_global"]
    G = 1 << g
    assert world is None or world == G
    shards = [np.zeros(1 << nl, dtype=complex) for _ in range(G)]
    shards[0][0] = 1.0
    vals = np.zeros(info["n_observables"])
    for kind, first, count, _ in info["segs"]:
        if kind == 1:  # EXCHANGE: top g local bits <-> rank bits
            blk = 1 << (nl - g)
            new = [np.empty_like(s) for s in shards]
            for s in range(G):
                for v in range(G):
                    new[v][s * blk:(s + 1) * blk] = shards[s][v * blk:(v + 1) * blk]
            shards = new
        elif kind == 2:
            for r in range(G):
                zexp(info, shards[r], r, first, count, vals)
        else:
            for r in range(G):
                apply_sweeps(info, shards[r], r, first, count)
    return vals, shards


class EmulatorExecutor:
    """Drop-in for the GPU executor of ml_qem_b200.statevector.ShardedStatevector: runs the local
    segments of one rank on CPU torch tensors, so the distributed orchestration (segment loop,
    all_to_all_single exchange, all_reduce of the values) can be tested under gloo."""

    device = "cpu"

    def prepare(self, program):
        self.info = program.info

    def init_state(self, state, rank):
        state.zero_()
        if rank == 0:
            state[0] = 1.0

    def run_segment(self, program, seg, state, rank, obs):
        kind, first, count, _ = (int(x) for x in program.info["segs"][seg])
        psi = state.numpy()
        if kind == 0:
            apply_sweeps(program.info, psi, rank, first, count)
        else:
            v = obs.numpy()
            zexp(program.info, psi, rank, first, count, v)
