"""Lock-step emulation of ``cond_factor_group_kernel`` (csrc/condense.cuh): G lanes per edge share the banded LU
with partial pivoting and the solves for the columns of Y_e.  Every phase between two ``__syncwarp()`` is replayed
with all lanes' reads and writes recorded: no lane may write what another lane reads or writes in the same phase,
and pivots / factors must equal those of the sequential scheme (``cond_band_lu`` / ``cond_band_solve``) bit for bit.
Test infrastructure: it restates the kernel's index arithmetic, it is not a CPU path of the product."""

import numpy as np
import pytest

from networks_fenicsx_b200 import condense


class Phase:
    """The accesses of all lanes between two warp barriers: writes become visible at ``commit``."""

    def __init__(self, mem):
        self.mem, self.writes, self.reads = mem, {}, {}

    def read(self, lane, addr):
        self.reads.setdefault(addr, set()).add(lane)
        return self.mem[addr]

    def write(self, lane, addr, value):
        assert addr not in self.writes or self.writes[addr][0] == lane, ("two lanes write", addr)
        self.writes[addr] = (lane, value)

    def commit(self):
        for addr, (lane, value) in self.writes.items():
            others = self.reads.get(addr, set()) - {lane}
            assert not others, ("a lane reads what another lane writes in the same phase", addr, lane, others)
            self.mem[addr] = value


def seq_lu(K, n, kl):
    """cond_band_lu on a dense array (dgbtf2)."""
    A = K.copy()
    piv = np.zeros(n, dtype=int)
    ju = 0
    for j in range(n):
        km = min(kl, n - 1 - j)
        jp = int(np.argmax(np.abs(A[j:j + km + 1, j])))
        piv[j] = j + jp
        ju = max(ju, min(j + jp + kl, n - 1))
        if jp:
            A[[j, j + jp], j:ju + 1] = A[[j + jp, j], j:ju + 1]
        A[j + 1:j + km + 1, j] *= 1.0 / A[j, j]
        for col in range(j + 1, ju + 1):
            A[j + 1:j + km + 1, col] -= A[j + 1:j + km + 1, j] * A[j, col]
    return A, piv


def seq_solve(LU, piv, n, kl, rhs):
    """cond_band_solve (dgbtrs, no transpose)."""
    b = rhs.copy()
    kv = 2 * kl
    for j in range(n):
        l = piv[j]
        bj = b[l]
        if l != j:
            b[l] = b[j]
            b[j] = bj
        lm = min(kl, n - 1 - j)
        b[j + 1:j + lm + 1] -= LU[j + 1:j + lm + 1, j] * bj
    for j in range(n - 1, -1, -1):
        bj = b[j] / LU[j, j]
        b[j] = bj
        lo = max(0, j - kv)
        b[lo:j] -= LU[lo:j, j] * bj
    return b


def group_factor(K, Cm, n, n_max, kl, G, active):
    """The kernel, phase by phase: band element (i, j) lives at ``(kv + i - j) * n_max + j``; lane = column of
    Y + 4 * helper in the solves."""
    kv, ldab = 2 * kl, 3 * kl + 1
    mem = {("A", k): 0.0 for k in range(ldab * n_max)}

    def B(i, j):
        return ("A", (kv + i - j) * n_max + j)

    for i, j in zip(*np.nonzero(K)):
        assert abs(i - j) <= kl
        mem[B(i, j)] = K[i, j]
    piv = [0] * n_max
    ju = 0
    for j in range(n_max):  # ---- LU: pivot search (all lanes, redundantly) | swap | scale | rank-1 update
        act = j < n
        km = jp = 0
        if act:
            km = min(kl, n - 1 - j)
            best = abs(mem[B(j, j)])
            for i in range(1, km + 1):
                a = abs(mem[B(j + i, j)])
                if a > best:
                    best, jp = a, i
            piv[j] = j + jp
            ju = max(ju, min(j + jp + kl, n - 1))
        ph = Phase(mem)
        if act and jp:
            for lane in range(G):
                for col in range(j + lane, ju + 1, G):
                    a, b = ph.read(lane, B(j, col)), ph.read(lane, B(j + jp, col))
                    ph.write(lane, B(j, col), b)
                    ph.write(lane, B(j + jp, col), a)
        ph.commit()
        ph = Phase(mem)
        if act:
            for lane in range(G):
                inv = 1.0 / ph.read(lane, B(j, j))
                for i in range(1 + lane, km + 1, G):
                    ph.write(lane, B(j + i, j), ph.read(lane, B(j + i, j)) * inv)
        ph.commit()
        ph = Phase(mem)
        if act:
            total = km * (ju - j)
            for lane in range(G):
                for idx in range(lane, total, G):
                    i, col = 1 + idx % km, j + 1 + idx // km
                    ph.write(lane, B(j + i, col),
                             ph.read(lane, B(j + i, col)) - ph.read(lane, B(j + i, j)) * ph.read(lane, B(j, col)))
        ph.commit()
    H = G // 4
    for s in range(4):
        for k in range(n_max):
            mem[("Y", s * n_max + k)] = Cm[k, s] if k < n else 0.0
    for j in range(n_max):  # ---- forward: read | swap (helper 0) | update (helpers share the rows)
        act = j < n
        ph, held = Phase(mem), {}
        if act:
            l = piv[j]
            for lane in range(G):
                s = lane % 4
                if active[s]:
                    held[lane] = (ph.read(lane, ("Y", s * n_max + l)), ph.read(lane, ("Y", s * n_max + j)))
        ph.commit()
        ph = Phase(mem)
        if act and l != j:
            for lane in range(G):
                s, h = lane % 4, lane // 4
                if active[s] and h == 0:
                    ph.write(lane, ("Y", s * n_max + l), held[lane][1])
                    ph.write(lane, ("Y", s * n_max + j), held[lane][0])
        ph.commit()
        ph = Phase(mem)
        if act:
            lm = min(kl, n - 1 - j)
            for lane in range(G):
                s, h = lane % 4, lane // 4
                if active[s]:
                    for i in range(1 + h, lm + 1, H):
                        addr = ("Y", s * n_max + j + i)
                        ph.write(lane, addr, ph.read(lane, addr) - ph.read(lane, B(j + i, j)) * held[lane][0])
        ph.commit()
    for j in range(n_max - 1, -1, -1):  # ---- backward: read + divide | store (helper 0) and update
        act = j < n
        ph, held = Phase(mem), {}
        if act:
            for lane in range(G):
                s = lane % 4
                if active[s]:
                    held[lane] = ph.read(lane, ("Y", s * n_max + j)) / ph.read(lane, B(j, j))
        ph.commit()
        ph = Phase(mem)
        if act:
            for lane in range(G):
                s, h = lane % 4, lane // 4
                if not active[s]:
                    continue
                if h == 0:
                    ph.write(lane, ("Y", s * n_max + j), held[lane])
                for i in range(max(0, j - kv) + h, j, H):
                    addr = ("Y", s * n_max + i)
                    ph.write(lane, addr, ph.read(lane, addr) - ph.read(lane, B(i, j)) * held[lane])
        ph.commit()
    LU = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if (0 <= i - j <= kl) or (0 < j - i <= kv):
                LU[i, j] = mem[B(i, j)]
    Y = np.array([[mem[("Y", s * n_max + k)] for s in range(4)] for k in range(n)])
    return LU, piv[:n], Y


@pytest.mark.parametrize("fd,pd,N", [(2, 1, 4), (3, 2, 4), (2, 0, 3), (2, 1, 1), (4, 3, 2)])
def test_lane_phases_are_hazard_free_and_reproduce_the_sequential_factors(fd, pd, N):
    rng = np.random.default_rng(10 * fd + pd)
    for typ in range(4):
        t = condense.edge_type_tables(N, fd, pd, bool(typ & 1), bool(typ & 2))
        n = t.n
        n_max = n + 1  # the kernel loops to the largest local size of all edge types
        kl = int(np.abs(t.k_row - t.k_col).max())
        rh = rng.uniform(0.5, 2, N)
        K = np.zeros((n, n))
        np.add.at(K, (t.k_row, t.k_col), t.k_coef * np.where(t.k_cell >= 0, rh[np.maximum(t.k_cell, 0)], 1.0))
        Cm = np.zeros((n, 4))
        np.add.at(Cm, (t.c_row, t.c_slot), t.c_coef)
        active = [bool(typ & 1) and pd >= 1, bool(typ & 1), bool(typ & 2) and pd >= 1, bool(typ & 2)]
        LUs, pivs = seq_lu(K, n, kl)
        mask = np.array([[(0 <= i - j <= kl) or (0 < j - i <= 2 * kl) for j in range(n)] for i in range(n)])
        for G in (8, 16, 32):
            LU, piv, Y = group_factor(K, Cm, n, n_max, kl, G, active)
            assert list(piv) == list(pivs)
            assert np.array_equal(LU[mask], LUs[mask]), (typ, G)
            for s in range(4):
                if active[s]:
                    assert np.array_equal(Y[:, s], seq_solve(LUs, pivs, n, kl, Cm[:, s])), (typ, G, s)
                    assert np.allclose(K @ Y[:, s], Cm[:, s], atol=1e-10)
                else:
                    assert not Y[:, s].any()
