"""numpy model of the quality pipeline's scan algebra (test infrastructure).

Mirrors demodulator_b200/csrc/fmd_quality.cuh step for step -- thread-level zero-state runs, the
CTA-level affine scan with the host-built power tables, the tile-level scan, the zero-input
response corrections and the decimation bookkeeping -- in float64 on the float32 tables that
libfmdemod.so builds (fmq_test_build).  It exists so that the algebra and the tables can be
checked against the serial oracle on a machine without a GPU; it is not a product path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

THREADS, MAXD = 256, 6


class _Consts(C.Structure):
    _fields_ = [("L", C.c_int), ("D", C.c_int), ("sos", C.c_float * 5 * 3), ("lvl", C.c_float * 36 * 5),
                ("warp", C.c_float * 36), ("tile", C.c_float * 36)]


class Filter:
    def __init__(self, lib, sos: np.ndarray, T: int) -> None:
        """T = samples per thread (64, or 128 when the decimation factor is large); tile = 256 T"""
        self.T, self.TILE = T, T * THREADS
        TILE = self.TILE
        sos = np.ascontiguousarray(sos, np.float32)
        L = sos.shape[0]
        D = 2 * L
        nbytes = C.c_uint(0)
        sizes = (C.c_uint * 5)()
        lib.fmq_test_sizes.argtypes = [C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        lib.fmq_test_build.argtypes = ([C.c_void_p, C.c_uint, C.c_uint, C.POINTER(_Consts), C.POINTER(_Consts)] +
                                       [C.c_void_p] * 6 + [C.POINTER(C.c_uint)])
        lib.fmq_test_sizes(L, T, C.byref(nbytes), sizes)
        assert nbytes.value == C.sizeof(_Consts), (nbytes.value, C.sizeof(_Consts))
        assert sizes[2] == T * D and sizes[3] == TILE * D
        self.c, self.ct = _Consts(), _Consts()
        lane = np.zeros(sizes[0], np.float32)
        apow = np.zeros(sizes[1], np.float32)
        ck = np.zeros(sizes[2], np.float32)
        ckn = np.zeros(sizes[3], np.float32)
        lane_tile = np.zeros(sizes[4], np.float32)
        abcd = np.zeros(D * D + 2 * D + 1, np.float64)
        zl = C.c_uint(0)
        rc = lib.fmq_test_build(sos.ctypes.data, L, T, C.byref(self.c), C.byref(self.ct), lane.ctypes.data,
                                apow.ctypes.data, ck.ctypes.data, ckn.ctypes.data, lane_tile.ctypes.data,
                                abcd.ctypes.data, C.byref(zl))
        assert rc == 0
        self.zir_len = int(zl.value)
        self.L, self.D = L, D
        self.lvl_tile = [np.array(self.ct.lvl[j], np.float64)[:D * D].reshape(D, D) for j in range(5)]
        self.warp_tile = np.array(self.ct.warp, np.float64)[:D * D].reshape(D, D)
        self.lane_tile = lane_tile.astype(np.float64).reshape(32, D, D)
        self.sos = np.array(self.c.sos, np.float64)[:L]  # b0 b1 b2 a1 a2
        self.lvl = [np.array(self.c.lvl[j], np.float64)[:D * D].reshape(D, D) for j in range(5)]
        self.warp = np.array(self.c.warp, np.float64)[:D * D].reshape(D, D)
        self.tile = np.array(self.c.tile, np.float64)[:D * D].reshape(D, D)
        self.lane = lane.astype(np.float64).reshape(32, D, D)
        self.apow = apow.astype(np.float64).reshape(T + 1, D, D)
        self.ck = ck.astype(np.float64).reshape(T, D)
        self.ckn = ckn.astype(np.float64).reshape(TILE, D)
        self.A = abcd[:D * D].reshape(D, D)

    def step(self, q: np.ndarray, u: np.ndarray) -> np.ndarray:
        """one sample for every row of q (rows = threads); q is updated in place"""
        for i in range(self.L):
            b0, b1, b2, a1, a2 = self.sos[i]
            w = b0 * u + q[:, 2 * i]
            q[:, 2 * i] = b1 * u - a1 * w + q[:, 2 * i + 1]
            q[:, 2 * i + 1] = b2 * u - a2 * w
            u = w
        return u

    def matpow(self, k: int) -> np.ndarray:
        return np.linalg.matrix_power(self.A, k)


def block_scan(f: "Filter", e: np.ndarray, carry: np.ndarray) -> np.ndarray:
    """e: [ntiles, 256, D] zero-state end states; carry: [ntiles, D] -> state at each thread's start"""
    nt = e.shape[0]
    v = e.reshape(nt, 8, 32, f.D).copy()
    for j in range(5):
        off = 1 << j
        up = np.zeros_like(v)
        up[:, :, off:] = v[:, :, :-off]
        add = np.einsum("ik,twlk->twli", f.lvl[j], up)
        add[:, :, :off] = 0.0
        v = v + add
    tot = v[:, :, 31]  # [nt, 8, D]
    wp = np.zeros((nt, 8, f.D))
    cur = carry.copy()
    for w in range(8):
        wp[:, w] = cur
        cur = cur @ f.warp.T + tot[:, w]
    ex = np.zeros_like(v)
    ex[:, :, 1:] = v[:, :, :-1]
    ex += np.einsum("lik,twk->twli", f.lane, wp)
    return ex.reshape(nt, THREADS, f.D)


def warp_scan(lvl, v: np.ndarray) -> np.ndarray:
    """fmq_warp_scan over axis 0 (32 lanes) of v[32, D]"""
    v = v.copy()
    for j in range(5):
        off = 1 << j
        up = np.zeros_like(v)
        up[off:] = v[:-off]
        add = up @ lvl[j].T
        add[:off] = 0.0
        v = v + add
    return v


def scan_tiles(f: "Filter", E: np.ndarray, carry: np.ndarray, len_last: int):
    """fmq_scan_tiles_kernel: 32 warps, each a run of `runlen` tiles walked 32 tiles at a time with
    the tile-level tables; returns (state at the start of every tile, carry out)"""
    nt, D = E.shape
    TILE = f.TILE
    nwarps = 32
    runlen = ((nt + nwarps - 1) // nwarps + 31) & ~31
    prun = f.matpow(TILE * runlen)
    plast = f.matpow(len_last)
    runs = [(min(w * runlen, nt), min(min(w * runlen, nt) + runlen, nt)) for w in range(nwarps)]
    run_e = np.zeros((nwarps, D))
    for w, (j0, j1) in enumerate(runs):
        cw = np.zeros(D)
        for base in range(j0, j1, 32):
            v = np.zeros((32, D))
            m = min(32, j1 - base)
            v[:m] = E[base:base + m]
            cw = f.warp_tile @ cw + warp_scan(f.lvl_tile, v)[31]
        run_e[w] = cw
    Q = np.zeros_like(E)
    out = carry.copy()
    for w, (j0, j1) in enumerate(runs):
        c = carry.copy()
        for w2 in range(w):
            c = prun @ c + run_e[w2]
        for base in range(j0, j1, 32):
            v = np.zeros((32, D))
            m = min(32, j1 - base)
            v[:m] = E[base:base + m]
            inc = warp_scan(f.lvl_tile, v)
            ex = np.zeros((32, D))
            ex[1:] = inc[:-1]
            q = ex + np.einsum("lik,k->li", f.lane_tile, c)
            Q[base:base + m] = q[:m]
            if base + m == nt:
                out = plast @ q[m - 1] + E[nt - 1]
            c = f.warp_tile @ c + inc[31]
    return Q, out


def kept_below(pos, r0, M):
    pos = np.asarray(pos, np.int64)
    return np.where(pos <= r0, 0, (pos - r0 + M - 1) // M)


class Stream:
    """the host side of one quality context (fmdemod.cu: quality_launch) over the model kernels"""

    def __init__(self, lib, sos_in, sos_out, decim: int, T: int | None = None) -> None:
        T = T or (128 if decim >= 8 else 64)  # fmq_setup's choice
        self.T, self.TILE = T, T * THREADS
        self.fin = Filter(lib, sos_in, T) if sos_in is not None and len(sos_in) else None
        self.fout = Filter(lib, sos_out, T)
        self.M = decim
        self.pos = 0
        self.in_q = np.zeros((2, self.fin.D)) if self.fin else None
        self.out_q = np.zeros(self.fout.D)
        self.s_prev = np.zeros(2)

    def run(self, raw: np.ndarray) -> np.ndarray:
        n = raw.size // 2
        M = self.M
        T, TILE = self.T, self.TILE
        ntiles = (n + TILE - 1) // TILE
        r0 = (M - self.pos % M) % M
        n_out = int(kept_below(n, r0, M))
        len_last = n - (ntiles - 1) * TILE
        s = raw[:2 * n].astype(np.float64).reshape(n, 2) - 127.5
        pad = np.zeros((ntiles * TILE + 1, 2))
        pad[0] = self.s_prev
        pad[1:n + 1] = s  # pad[i + 1] = s[i], pad[0] = the sample before the call
        nthr = ntiles * THREADS
        F = np.arange(nthr) * T
        cnt = np.clip(n - F, 0, T)
        last = (cnt > 0) & (((np.arange(nthr) % THREADS) == THREADS - 1) | (F + T >= n))

        qi = None
        if self.fin:
            f = self.fin
            # fmq_in_zero_state: halo, then samples F .. F+cnt-2
            e = np.zeros((2, nthr, f.D))
            for lanei in range(2):
                for k in range(T):
                    act = k < cnt
                    u = pad[F + k, lanei]  # k = 0 is the halo (sample F-1)
                    q2 = e[lanei].copy()
                    f.step(q2, u)
                    e[lanei][act] = q2[act]
            # per-tile end states from a zero state, then the carry from tile to tile (on the GPU: the
            # decoupled look-back inside fmq_main_kernel -- the same recurrence q <- P q + E, resolved
            # in a different order)
            Ein = np.zeros((ntiles, 2, f.D))
            for lanei in range(2):
                ex = block_scan(f, e[lanei].reshape(ntiles, THREADS, f.D), np.zeros((ntiles, f.D))).reshape(nthr, f.D)
                idx = np.nonzero(last)[0]
                for t in idx:
                    Ein[t // THREADS, lanei] = f.apow[cnt[t]] @ ex[t] + e[lanei][t]
            Qin = np.zeros_like(Ein)
            for lanei in range(2):
                Qin[:, lanei], self.in_q[lanei] = scan_tiles(f, Ein[:, lanei], self.in_q[lanei], len_last)
            # main kernel, input part: state at F-1 of every thread
            qi = np.zeros((2, nthr, f.D))
            for lanei in range(2):
                qi[lanei] = block_scan(f, e[lanei].reshape(ntiles, THREADS, f.D), Qin[:, lanei]).reshape(nthr, f.D)

        # second pass
        fo = self.fout
        if self.fin:
            pr = self.fin.step(qi[0], pad[F, 0])
            pi = self.fin.step(qi[1], pad[F, 1])
        else:
            pr, pi = pad[F, 0].copy(), pad[F, 1].copy()
        eo = np.zeros((nthr, fo.D))
        vz = np.zeros((nthr, T))
        for k in range(T):
            act = k < cnt
            zr, zi = pad[F + k + 1, 0], pad[F + k + 1, 1]
            if self.fin:
                q0, q1 = qi[0].copy(), qi[1].copy()
                zr2, zi2 = self.fin.step(q0, zr), self.fin.step(q1, zi)
                qi[0][act], qi[1][act] = q0[act], q1[act]
                zr, zi = zr2, zi2
            re = zr * pr + zi * pi
            im = zi * pr - zr * pi
            x = np.arctan2(im, re)
            pr, pi = np.where(act, zr, pr), np.where(act, zi, pi)
            q2 = eo.copy()
            v = fo.step(q2, x)
            eo[act] = q2[act]
            vz[:, k] = v
        qo = block_scan(fo, eo.reshape(ntiles, THREADS, fo.D), np.zeros((ntiles, fo.D))).reshape(nthr, fo.D)
        Eout = np.zeros((ntiles, fo.D))
        for t in np.nonzero(last)[0]:
            Eout[t // THREADS] = fo.apow[cnt[t]] @ qo[t] + eo[t]
        # kept samples, thread-level correction
        o = np.arange(n_out)
        ks = r0 + o * M
        thr, k = ks // T, ks % T
        out = vz[thr, k] + np.einsum("oi,oi->o", fo.ck[k], qo[thr])
        # tile scan + tile-level correction
        Qout, self.out_q = scan_tiles(fo, Eout, self.out_q, len_last)
        zir = np.einsum("oi,oi->o", fo.ckn[ks % TILE], Qout[ks // TILE])
        out = out + np.where(ks % TILE < fo.zir_len, zir, 0.0)  # fmq_zir_kernel stops at zir_len
        self.s_prev = s[n - 1].copy()
        self.pos += n
        return out.astype(np.float32)
