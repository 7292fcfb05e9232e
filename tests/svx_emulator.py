"""numpy emulator of the statevector sweep program (test infrastructure): executes the segments of
an SvxProgram -- tile sweeps with the kernel's op kinds and index conventions, the EXCHANGE
all-to-all between simulated ranks, Z-type expectation sums -- so the planner (host C++) can be
checked against the oracle on a CPU-only box, including the amplitude-sharded case."""
import numpy as np

SVO_U1, SVO_X, SVO_U2, SVO_SWAP, SVO_D1, SVO_D2, SVO_R1, SVO_X1, SVO_DZZ = 1, 2, 3, 4, 5, 6, 7, 8, 9
SVF_COND, SVF_COND_VAL = 2, 4


SVS_GENERIC, SVS_X1, SVS_R1, SVS_U1, SVS_DIAG = 0, 1, 5, 9, 13


def svz12(j):
    """shared-memory swizzle of the statevector tile (program.h)."""
    return j ^ ((j >> 3) & 7) ^ ((j >> 6) & 7) ^ ((j >> 9) & 7)


def decode_block(info, sw):
    words = info["prog"]
    blk = words[2 * int(sw[0]): 2 * (int(sw[0]) + (int(sw[9]) & 0xffff))]  # high half: first-pass flag
    b = blk.view(np.uint8)
    n_passes = int(blk[:1].view(np.int32)[0])
    passes = []
    for p in range(n_passes):
        h = b[16 + 96 * p: 16 + 96 * (p + 1)]
        ops_q8, n_ops = int(h[:2].view(np.uint16)[0]), int(h[2:4].view(np.uint16)[0])
        ops = []
        for o in range(n_ops):
            r = b[8 * (ops_q8 + o): 8 * (ops_q8 + o + 1)]
            ops.append(dict(kind=int(r[0]), flags=int(r[1]), qa=int(r[2]), qb=int(r[3]),
                            off=int(r[4:6].view(np.uint16)[0]), cond_bit=int(r[6])))
        passes.append(dict(s=[int(x) for x in h[4:8]], sig=int(h[8]), n_pre=int(h[9]), needs_index=int(h[10]), flags=int(h[11]),
                           pp=[int(x) for x in h[12:16]], tb=[int(x) for x in h[16:24]],
                           cor=[int(x) for x in h[32:96].view(np.uint32)], ops=ops))
    return passes, blk.view(np.float64)


def check_pass_header(ph, slotpos, K):
    """Host-computed addressing of a pass: slots distinct and resident, positions consistent, the
    thread -> element map covers the tile exactly once, quarter warps are bank-conflict free (2^12
    tiles), and a fast signature really has the shape [diagonal][same-kind slot ops on 0..n-1][diagonal]."""
    s = ph["s"]
    assert len(set(s)) == 4 and all(0 <= x < K for x in s)
    assert ph["pp"] == [slotpos[x] for x in s]
    assert ph["cor"] == [16 * svz12(sum(((c >> i) & 1) << s[i] for i in range(4))) for c in range(16)]
    nt = K - 4
    tb = ph["tb"][:nt]
    assert sorted(tb + s) == list(range(K))
    tid = np.arange(1 << nt)
    j = np.zeros(1 << nt, dtype=np.int64)
    for k in range(nt):
        j |= ((tid >> k) & 1) << tb[k]
    seen = np.zeros(1 << K, dtype=np.int32)
    for c in ph["cor"]:
        np.add.at(seen, (16 * svz12(j) ^ c) // 16, 1)
    assert (seen == 1).all()
    if K == 12:
        a = 16 * svz12(j)
        for c in ph["cor"]:
            for q in range(0, 256, 8):
                assert len(set((((a[q:q + 8] ^ c) >> 4) & 7).tolist())) == 8
    assert ph["sig"] != SVS_GENERIC or ph["flags"] == 0  # only fast passes load / store directly
    if ph["sig"] != SVS_GENERIC:
        kinds = [o["kind"] for o in ph["ops"]]
        n_pre = ph["n_pre"]
        n_slot = 0 if ph["sig"] == SVS_DIAG else (ph["sig"] - 1) % 4 + 1
        want = {SVS_X1: SVO_X1, SVS_R1: SVO_R1, SVS_U1: SVO_U1}.get(ph["sig"] - (n_slot - 1) if n_slot else -1)
        assert all(k in (SVO_D1, SVO_D2, SVO_DZZ) for k in kinds[:n_pre] + kinds[n_pre + n_slot:])
        for i, o in enumerate(ph["ops"][n_pre:n_pre + n_slot]):
            assert o["kind"] == want and o["qa"] == i and not (o["flags"] & SVF_COND)
        # the register-resident part is what talks to global memory on a direct pass
        assert not (ph["flags"] & 1) or (n_pre == 0 and n_slot > 0)
        assert not (ph["flags"] & 2) or (len(kinds) == n_pre + n_slot and n_slot > 0)


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
        for ph in passes:
            check_pass_header(ph, slotpos, K)
            needs_index, ops, pp = ph["needs_index"], ph["ops"], ph["pp"]
            for op in ops:
                cond = np.ones(1 << nl, dtype=bool)
                if op["flags"] & SVF_COND:
                    assert needs_index
                    want = 1 if op["flags"] & SVF_COND_VAL else 0
                    cond = ((gi >> op["cond_bit"]) & 1) == want
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
                    pt = pp[op["qa"]]
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
                    p0, p1 = pp[op["qa"]], pp[op["qb"]]  # matrix index i_p0 + 2 i_p1
                    assert p0 != p1
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
