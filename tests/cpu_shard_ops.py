"""CPU stand-in for the three sharded-flow kernels (TEST INFRASTRUCTURE ONLY).

Implements the message / state contract of csrc/spx_pick.cu and csrc/spx_update.cu in
numpy so that simplex_method_solver_b200/parallel.py (shard ranges, all-gather layout,
ping-pong parity, termination) can be exercised under gloo without a GPU.  numpy
evaluates t*p, r*c, the difference and the quotient as separate IEEE operations, so
the arithmetic is the reference's (simplex.py:173-175).
"""
import ctypes

import numpy as np

from simplex_method_solver_b200 import _native as N

NONE = 0x7FFFFFFF
U64_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def _state(t):
    return N.SpxState.from_buffer(t.numpy())          # shares memory with the CPU tensor


class CpuShardOps:
    def import_shard(self, rows, function, A, b, n, m, col0, m_loc, ld):
        a = A.numpy()
        a[:] = 0.0
        a[:n, :m_loc] = rows[:, col0:col0 + m_loc]
        a[n, :m_loc] = function[col0:col0 + m_loc]
        b.numpy()[:n] = rows[:, m]

    def init_state(self, state, rowlab, collab, n, m, max_pivots):
        rowlab.numpy()[:] = np.arange(m)
        collab.numpy()[:n] = np.arange(m, m + n)
        st = _state(state)
        ctypes.memset(ctypes.addressof(st), 0, ctypes.sizeof(st))
        st.status, st.r, st.c, st.max_pivots, st.slot = N.PIVOT, -1, -1, max_pivots, 1
        st.hint_tag[0] = st.hint_tag[1] = -1

    def candidate(self, A, b, n, m_loc, ld, col0, rule, state, send):
        st = _state(state)
        if st.status != N.PIVOT:
            return
        a, bb = A.numpy(), b.numpy()[:n]
        neg = np.nonzero(bb < 0)[0]
        r1 = int(neg[0]) if len(neg) else -1
        line = a[r1, :m_loc] > 0 if r1 >= 0 else a[n, :m_loc] < 0
        hit = np.nonzero(line)[0]
        msg = send.numpy()
        hdr = msg[:4].view(np.uint64)
        hdr[2] = np.int64(r1).astype(np.uint64)
        hdr[3] = 0
        if len(hit) == 0:
            hdr[0] = hdr[1] = U64_MAX
            return
        j = int(hit[0])
        hdr[0] = 0
        hdr[1] = np.uint64(col0 + j)
        msg[4:4 + n + 1] = a[:, j]

    def select(self, gathered, world, b, n, rule, state, colbuf, flags=None, seq=0):
        st = _state(state)
        if st.status != N.PIVOT:
            return
        g = gathered.numpy()
        best, win = None, -1
        for k in range(world):
            hdr = g[k, :4].view(np.uint64)
            if hdr[1] == U64_MAX:
                continue
            key = (int(hdr[0]), int(hdr[1]))
            if best is None or key < best:
                best, win = key, k
        r1 = int(g[0, :4].view(np.int64)[2])
        if win < 0:
            st.status = N.INCORRECT if r1 >= 0 else N.OPTIMAL
            st.phase1 = int(r1 >= 0)
            return
        col = g[win, 4:4 + n + 1]
        colbuf.numpy()[:n + 1] = col
        bb = b.numpy()[:n]
        if r1 >= 0:
            r = r1
        else:
            # the reference's sequential scan, simplex.py:107-139
            r, first, mv = -1, True, 1.0
            for i in range(n):
                if col[i] == 0:
                    continue
                with np.errstate(all="ignore"):
                    v = bb[i] / col[i]
                if first:
                    mv, r, first = v, i, False
                elif (v == 0 and mv > 0) or (v < 0 <= mv) or (mv <= v < 0):
                    mv, r = v, i
            if first or mv > 0:
                st.status, st.phase1 = N.NOCONV, 0
                return
        st.phase1 = int(r1 >= 0)
        if st.npiv >= st.max_pivots:
            st.status = N.CAP
            return
        st.status, st.r, st.c, st.p = N.PIVOT, r, best[1], float(col[r])
        st.slot = (st.npiv + 1) & 1

    # ---- look-ahead halves: price pivot k+1 from table k -------------------------------
    @staticmethod
    def _next_local(a, col, r, p, cl, m_loc):
        """The local columns of the NEXT table (numpy, the update's arithmetic)."""
        with np.errstate(all="ignore"):
            new = (a[:, :m_loc] * p - a[r, :m_loc][None, :] * col[:, None]) / p
            new[r, :] = -a[r, :m_loc] / p
            if 0 <= cl < m_loc:
                new[:, cl] = col / p
                new[r, cl] = 1.0 / p
        return new

    def ahead_candidate(self, A, bin_, bout, n, m_loc, ld, col0, rule, state, colbuf, send):
        st = _state(state)
        msg = send.numpy()
        hdr = msg[:4].view(np.uint64)
        if st.status != N.PIVOT:
            hdr[0] = hdr[1] = hdr[2] = U64_MAX
            hdr[3] = 1
            return
        r, p = st.r, st.p
        col = colbuf.numpy()[:n + 1]
        bi = bin_.numpy()[:n]
        with np.errstate(all="ignore"):
            nb = (bi * p - bi[r] * col[:n]) / p
            nb[r] = -bi[r] / p
        bout.numpy()[:n] = nb
        new = self._next_local(A.numpy(), col, r, p, st.c - col0, m_loc)
        neg = np.nonzero(nb < 0)[0]
        r1 = int(neg[0]) if len(neg) else -1
        line = new[r1] > 0 if r1 >= 0 else new[n] < 0
        hit = np.nonzero(line)[0]
        hdr[2] = np.int64(r1).astype(np.uint64)
        hdr[3] = 0
        if len(hit) == 0:
            hdr[0] = hdr[1] = U64_MAX
            return
        j = int(hit[0])
        hdr[0] = 0
        hdr[1] = np.uint64(col0 + j)
        msg[4:4 + n + 1] = new[:, j]

    def ahead_select(self, gathered, world, bnext, n, state_cur, state_next, colbuf_next, flags=None, seq=0):
        cur, nxt = _state(state_cur), _state(state_next)
        g = gathered.numpy()
        if cur.status != N.PIVOT or g[0, :4].view(np.uint64)[3] != 0:
            ctypes.memmove(ctypes.addressof(nxt), ctypes.addressof(cur), ctypes.sizeof(cur))
            nxt.hint_tag[0] = nxt.hint_tag[1] = -1
            return
        ctypes.memmove(ctypes.addressof(nxt), ctypes.addressof(cur), ctypes.sizeof(cur))
        nxt.npiv = cur.npiv + 1
        nxt.status = N.PIVOT
        nxt.hint_tag[0] = nxt.hint_tag[1] = -1
        # the classic select on the next b, against the advanced pivot counter
        self.select(gathered, world, bnext, n, 0, state_next, colbuf_next)

    def update(self, Ain, Aout, bin_, bout, n, m_loc, ld, col0, state, colbuf, rowlab, collab, trace,
               ahead=False):
        st = _state(state)
        if st.status != N.PIVOT:
            return
        r, cg, p = st.r, st.c, st.p
        a = Ain.numpy()
        col = colbuf.numpy()[:n + 1]
        with np.errstate(all="ignore"):
            new = (a[:, :m_loc] * p - a[r, :m_loc][None, :] * col[:, None]) / p
            new[r, :] = -a[r, :m_loc] / p
            cl = cg - col0
            if 0 <= cl < m_loc:
                new[:, cl] = col / p
                new[r, cl] = 1.0 / p
            bi = bin_.numpy()[:n]
            nb = (bi * p - bi[r] * col[:n]) / p
            nb[r] = -bi[r] / p
        Aout.numpy()[:, :m_loc] = new
        rl, cl_ = rowlab.numpy(), collab.numpy()
        rl[cg], cl_[r] = cl_[r], rl[cg]
        if trace is not None:
            trace.numpy()[st.npiv] = (r, cg)
        if not ahead:                 # look-ahead: b and the state belong to the ahead halves
            bout.numpy()[:n] = nb
            st.npiv += 1
