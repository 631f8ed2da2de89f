"""CPU check of the panel schedule of ``solve_kernel_v3`` (``csrc/symbolic3.cpp``): the tables are executed in NumPy in
exactly the phase order of the kernel (early updates with partial blocks -> solves of the previous panel -> fold + late
update -> diagonal chain, then the backward substitution on the stored factor) and the result is compared with a dense solve.
What the kernel adds on top of this is layout and synchronisation only, so a schedule bug shows up here, without a GPU."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_kernels import _grid3d_couplings, _random_reduced_system

W = 16


def _schedule(sysd):
    from pylrbms_b200._lib import Symbolic, load_library, ptr
    lib = load_library()
    sym = Symbolic(sysd['sizes'], [b[0] for b in sysd['blocks']], [b[1] for b in sysd['blocks']])
    s3 = C.c_void_p()
    assert lib.lrbms_symbolic3_create(sym.s, C.byref(s3)) == 0

    def info(k):
        out = C.c_int64()
        assert lib.lrbms_symbolic3_info(s3, k, C.byref(out)) == 0
        return out.value

    def get(which, n):
        out = np.zeros(n, dtype=np.int32)
        assert lib.lrbms_symbolic3_get(s3, which, ptr(out), n) == n
        return out
    T = dict(ok=info(0), np=info(1), n_win=info(2), acc_rows=info(3), n_partial=info(4), n_tiles=info(5), flops=info(6),
             own_words=info(7), pan_words=info(8), n_steps=info(9), ntc=sym.n_tile_cols, n_pad=sym.n_pad, flops_v2=sym.flops)
    if T['ok']:
        T['col_ptr'] = get(0, T['ntc'] + 1)
        T['row_idx'] = get(1, T['n_tiles'])
        T['a_map'] = get(2, T['n_tiles'])
        T['own'] = get(4, (T['np'] + 1) * W * T['own_words']).reshape(T['np'] + 1, W, T['own_words'])
        T['pan'] = get(5, T['np'] * T['pan_words']).reshape(T['np'], T['pan_words'])
        T['steps'] = get(6, 4 * T['n_steps']).reshape(-1, 4)
    if not T['ok']:
        T['why'] = ''.join(chr(c) for c in get(7, info(10)))
    lib.lrbms_symbolic3_destroy(s3)
    return T


def _own(rec):
    """csrc/symbolic3.h V3Own"""
    it = iter(range(len(rec)))
    take = lambda n: [int(rec[next(it)]) for _ in range(n)]
    o = dict(row=take(2), acc=take(2), prev=take(2))
    o['wprev'] = [take(2), take(2)]
    o['gprev'] = [take(2), take(2)]
    o['amap'] = [take(2), take(2)]
    o['exists'] = [take(2), take(2)]
    o['fold'] = take(4)
    o['n_chunks'] = take(1)[0]
    o['chunk_step'], o['chunk_n'], o['chunk_dest'], o['chunk_kind'] = take(4), take(4), take(4), take(4)
    return o


def _pan(rec):
    """csrc/symbolic3.h V3Panel"""
    it = iter(range(len(rec)))
    take = lambda n: [int(rec[next(it)]) for _ in range(n)]
    P = dict(zip(('c0', 'c1', 'g_d00', 'g_d10', 'g_d11', 'w_d10', 'a_d00', 'a_d10', 'a_d11'), take(9)))
    P['fold'] = take(4)
    P['head_prev'] = take(1)[0]
    P['head_exists'] = take(2)
    P['head_acc'] = take(2)
    P['head_w'] = [take(2), take(2)]
    P['head_g'] = [take(2), take(2)]
    P['acc_rows'] = take(2)
    P['step0'], P['n_steps'], _ = take(3)
    return P


def emulate(T, A, f):
    """A: dense SPD (n_pad x n_pad, identity on the padding), f: (n_pad,).  Returns u."""
    npan, Z = T['np'], T['n_win']
    tile = lambda I, c: A[8 * I:8 * I + 8, 8 * c:8 * c + 8]
    win = np.full((Z + 1, 8, 8), np.nan)
    win[Z] = 0.0
    L = {}
    AR = T['acc_rows']
    accbuf = np.full((2 * AR + 2, 8, 8), np.nan)
    part = np.full((max(1, T['n_partial']), 2, 2, 8, 8), np.nan)
    sx = np.full(T['n_pad'], np.nan)
    sW = [None, None]
    steps = T['steps']
    for p in range(-1, npan):
        q = p + 1
        owns = [_own(T['own'][q, w]) for w in range(W)]
        reg, X = {}, {}
        # ---- phase A: early updates of panel q
        if q < npan:
            t0, t1 = 2 * q, 2 * q + 1
            for w in range(1, W):
                o = owns[w]
                for k in range(o['n_chunks']):
                    dest, kind = o['chunk_dest'][k], o['chunk_kind'][k]
                    if dest == -2:
                        continue
                    acc = np.zeros((2, 2, 8, 8))
                    if dest == -1:
                        assert k == 0
                        if o['row'][0] == -2:
                            for c in range(2):
                                acc[0, c, 0, :] = f[8 * (t0 + c):8 * (t0 + c) + 8]
                        else:
                            for r in range(2):
                                for c in range(2):
                                    if o['row'][r] >= 0:
                                        blk = tile(o['row'][r], t0 + c)
                                        if o['amap'][r][c] < 0:
                                            assert not blk.any()
                                        acc[r, c] = blk
                    for s in range(o['chunk_step'][k], o['chunk_step'][k] + o['chunk_n'][k]):
                        a0, a1, b0, b1 = (int(x) for x in steps[s])
                        if kind == 1:
                            y = np.zeros((8, 8)); y[0] = sx[8 * a0:8 * a0 + 8]
                            assert not np.isnan(y).any()
                            acc[0, 0] -= y @ win[b0].T
                            acc[0, 1] -= y @ win[b1].T
                        else:
                            for r, a in enumerate((a0, a1)):
                                for c, b in enumerate((b0, b1)):
                                    acc[r, c] -= win[a] @ win[b].T
                    assert not np.isnan(acc).any(), 'an early update read a window tile that is not live'
                    if dest == -1:
                        reg[w] = acc
                    else:
                        part[dest] = acc
        # ---- phase B1: solves of panel p (own rows by the update warps, head rows by the chain warp)
        if p >= 0:
            W00, L10, W11 = sW[p & 1]

            def solve(C0, C1):
                X0 = C0 @ W00.T
                X1 = (C1 - X0 @ L10.T) @ W11.T
                return X0, X1
            for w in range(1, W):
                o = owns[w]
                if o['row'][0] == -2:
                    X0, X1 = solve(accbuf[2 * AR], accbuf[2 * AR + 1])
                    sx[16 * p:16 * p + 8], sx[16 * p + 8:16 * p + 16] = X0[0], X1[0]
                    X[w] = [[X0, X1], [np.zeros((8, 8))] * 2]
                    continue
                X[w] = [[np.zeros((8, 8))] * 2, [np.zeros((8, 8))] * 2]
                for r in range(2):
                    if o['row'][r] >= 0 and o['prev'][r]:
                        X0, X1 = solve(accbuf[2 * o['acc'][r]], accbuf[2 * o['acc'][r] + 1])
                        assert not np.isnan(X0).any() and not np.isnan(X1).any()
                        for c, Xc in enumerate((X0, X1)):
                            win[o['wprev'][r][c]] = Xc
                            L[o['gprev'][r][c]] = Xc
                        X[w][r] = [X0, X1]
        H = [[np.zeros((8, 8))] * 2, [np.zeros((8, 8))] * 2]
        if q < npan:
            P = _pan(T['pan'][q])
            if p >= 0:
                for r in range(2):
                    if P['head_exists'][r]:
                        X0, X1 = solve(accbuf[2 * P['head_acc'][r]], accbuf[2 * P['head_acc'][r] + 1])
                        for c, Xc in enumerate((X0, X1)):
                            win[P['head_w'][r][c]] = Xc
                            L[P['head_g'][r][c]] = Xc
                        H[r] = [X0, X1]
            D = np.zeros((2, 2, 8, 8))
            D[0, 0], D[1, 0], D[1, 1] = tile(P['c0'], P['c0']), tile(P['c1'], P['c0']), tile(P['c1'], P['c1'])
            for e in P['fold']:
                if e >= 0:
                    D += part[e]
        # ---- phase B2: fold + late update of panel q's blocks; diagonal chain of panel q
        if q <= npan:
            for w in range(1, W):
                o = owns[w]
                if w not in reg:
                    continue
                acc = reg[w].copy()
                for e in o['fold']:
                    if e >= 0:
                        assert not np.isnan(part[e]).any()
                        acc += part[e]
                if p >= 0:
                    for r in range(2):
                        for c in range(2):
                            for s in range(2):
                                acc[r, c] -= X[w][r][s] @ H[c][s].T
                if o['row'][0] == -2:
                    accbuf[2 * AR], accbuf[2 * AR + 1] = acc[0, 0], acc[0, 1]
                else:
                    for r in range(2):
                        if o['row'][r] >= 0:
                            accbuf[2 * o['acc'][r]], accbuf[2 * o['acc'][r] + 1] = acc[r, 0], acc[r, 1]
        if q < npan:
            for s in range(2):
                D[0, 0] -= H[0][s] @ H[0][s].T
                D[1, 0] -= H[1][s] @ H[0][s].T
                D[1, 1] -= H[1][s] @ H[1][s].T
            L00 = np.linalg.cholesky(D[0, 0])
            W00 = np.linalg.inv(L00)
            L10 = D[1, 0] @ W00.T
            L11 = np.linalg.cholesky(D[1, 1] - L10 @ L10.T)
            W11 = np.linalg.inv(L11)
            sW[q & 1] = (W00, L10, W11)
            L[P['g_d00']], L[P['g_d10']], L[P['g_d11']] = W00, L10, W11
    # ---- backward substitution on the stored factor (closed pattern; diagonal slots hold the inverses)
    u = sx.copy()
    cp, ri = T['col_ptr'], T['row_idx']
    assert len(L) == T['n_tiles'], 'not every tile of the factor was stored'
    for J in range(T['ntc'] - 1, -1, -1):
        v = u[8 * J:8 * J + 8].copy()
        for s in range(cp[J] + 1, cp[J + 1]):
            I = ri[s]
            v -= L[s].T @ u[8 * I:8 * I + 8]
        u[8 * J:8 * J + 8] = L[cp[J]].T @ v
    return u


CASES = [
    ('2x2 N8', 2, 2, [8] * 4, None),
    ('4x4 N8', 4, 4, [8] * 16, None),
    ('3x2 ragged', 3, 2, [3, 11, 7, 20, 1, 6], None),
    ('8x8 N20 (C2 shape)', 8, 8, [20] * 64, None),
    ('2x2x2 N16', 8, 1, [16] * 8, _grid3d_couplings(2, 2, 2)),
]


@pytest.mark.parametrize('name,sx,sy,sizes,couplings', CASES, ids=[c[0] for c in CASES])
def test_panel_schedule_reproduces_dense_solve(name, sx, sy, sizes, couplings):
    rng = np.random.default_rng(3)
    sysd = _random_reduced_system(rng, sx, sy, sizes, couplings=couplings)
    T = _schedule(sysd)
    if sum(sizes) % 16:
        assert not T['ok']              # odd number of tile columns: the one-column kernel handles it
        return
    assert T['ok'], (name, T.get('why'))
    n, n_pad = sysd['n'], T['n_pad']
    for mu in (0.15, 0.9):
        A = np.eye(n_pad)
        A[:n, :n] = sysd['dense'][0] + mu * sysd['dense'][1]
        f = np.zeros(n_pad)
        f[:n] = 1.7 * sysd['rhs'][0]
        u = emulate(T, A, f)
        ref = np.linalg.solve(A, f)
        assert np.abs(u - ref).max() <= 1e-11 * np.abs(ref).max()
    if name.startswith('8x8'):
        # the padding of the 2 x 2 blocks costs little on the band of the benchmark configuration
        assert T['flops'] <= 1.25 * T['flops_v2'], (T['flops'], T['flops_v2'])
        print('C2 shape: window tiles', T['n_win'], 'partial blocks', T['n_partial'], 'acc rows', T['acc_rows'],
              'flops v3 / v2', T['flops'] / T['flops_v2'])
