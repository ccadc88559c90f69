"""numpy model of the thread-level algorithm of csrc/shu_fft.cu (development aid: validates the index arithmetic -- thread
mapping, register rotations folded into the twiddles, row pairing, packed real columns, Hermitian handling -- against
numpy.fft before it is transcribed to CUDA).  python tools/proto/shu_fft_proto.py"""
import numpy as np


def dft8(v, sign):
    """v [..., 8] complex -> natural-order DFT-8 along the last axis, exp(sign*2*pi*i*nk/8)."""
    n = np.arange(8)
    w = np.exp(sign * 2j * np.pi * np.outer(n, n) / 8)
    return v @ w


def dftm(v, sign):
    m = v.shape[-1]
    n = np.arange(m)
    return v @ np.exp(sign * 2j * np.pi * np.outer(n, n) / m)


def fft_two_stage(v, t, M, sign, rin, rout, scale=1.0):
    """v [threads, 8]: thread t of an FFT group of M threads (L = 8*M) holds x[t + M*((a + rin) & 7)] in register a.
    Returns regs [threads, 8]: register q holds X[t + M*((q + rout') ...)] -- see below.  Groups are consecutive threads."""
    L = 8 * M
    nthr = v.shape[0]
    c = np.arange(8)
    Y = dft8(v, sign)                                     # over a' -> c ; = w8^{-rin c} * true
    if M == 1:
        # single thread: rotation un-done by twiddle; no second stage.  register c holds X[c]
        tw = np.exp(sign * 2j * np.pi * ((rin[:, None] * c[None, :]) % 8) / 8)
        return Y * tw * scale
    # stage-2 input twiddle: w_L^{t c} * w8^{rin c} * w_M^{t rout}
    expo = (t[:, None] * c[None, :]) / L + (rin[:, None] * c[None, :]) / 8 + (t * rout)[:, None] / M
    Y = Y * np.exp(sign * 2j * np.pi * expo) * scale
    # exchange inside each group: thread tt gets, for its c-values c = tt + M*e (e < 8/M), all b = 0..M-1
    out = np.zeros_like(Y)
    E = 8 // M
    for g0 in range(0, nthr, M):
        blk = Y[g0:g0 + M]                                # [b, c]
        for tt in range(M):
            for e in range(E):
                cc = tt + M * e
                u = blk[:, cc]                            # over b
                X = dftm(u, sign)                         # over b -> d (rotated by rout: X'[d] = X[d + rout])
                for d in range(M):
                    out[g0 + tt, e + E * d] = X[d]
    # register q = e + E*d holds X[cc + 8*((d + rout) % M)] = X[tt + M*e + 8*((d+rout)%M)]
    return out


def out_index(t, M, q, rout):
    E = 8 // M
    e, d = q % E, q // E
    return t + M * e + 8 * ((d + rout) % M) if M > 1 else q


def forward_plane(x):
    R = 64
    tid = np.arange(256)
    warp, lane = tid >> 5, tid & 31
    fl, t = lane >> 3, lane & 7
    # ---- row pass: FFT p packs rows p and p+32
    p = (warp & 3) + 4 * (fl & 1) + 8 * (fl >> 1) + 16 * (warp >> 2)
    assert sorted(set(p)) == list(range(32))
    rin = fl
    v = np.zeros((256, 8), complex)
    for a in range(8):
        col = 8 * ((a + rin) & 7) + t
        v[:, a] = x[p, col] + 1j * x[p + 32, col]
    rout = np.zeros(256, int)
    regs = fft_two_stage(v, t, 8, -1, rin, rout)
    Zs = np.zeros((32, 64), complex)
    for q in range(8):
        k = np.array([out_index(t[i], 8, q, rout[i]) for i in range(256)])
        Zs[p, k] = regs[:, q]
    # ---- column pass: FFT k (k = 0: packed kx 0 / 32)
    k = 4 * warp + fl
    v = np.zeros((256, 8), complex)
    for a in range(4):
        pp = t + 8 * a
        za = Zs[pp, k]
        zb = Zs[pp, np.where(k == 0, 32, (64 - k) & 63)]
        A = np.where(k == 0, za.real + 1j * zb.real, 0.5 * ((za.real + zb.real) + 1j * (za.imag - zb.imag)))
        B = np.where(k == 0, za.imag + 1j * zb.imag, 0.5 * ((za.imag + zb.imag) + 1j * (zb.real - za.real)))
        v[:, a] = A
        v[:, a + 4] = B
    rin = np.zeros(256, int)
    rout = fl
    regs = fft_two_stage(v, t, 8, -1, rin, rout, scale=1.0 / 4096)
    out = np.zeros((33, 64), complex)        # [kx][s]
    ky = np.zeros((256, 8), int)
    for q in range(8):
        ky[:, q] = [out_index(t[i], 8, q, rout[i]) for i in range(256)]
    for i in range(256):
        for q in range(8):
            s = (ky[i, q] - 33) & 63
            if k[i] != 0:
                out[k[i], s] = regs[i, q]
    # packed column: threads 0..7 (k == 0, rout == 0): C[ky], partner C[-ky]
    C = np.zeros(64, complex)
    for i in range(8):
        for q in range(8):
            C[ky[i, q]] = regs[i, q]
    for i in range(8):
        for q in range(8):
            kk = ky[i, q]
            cm = np.conj(C[(-kk) % 64])
            s = (kk - 33) & 63
            out[0, s] = 0.5 * (C[kk] + cm)
            out[32, s] = -0.5j * (C[kk] - cm)
    return out


def inverse_band(spec, gauss, R, r):
    """spec [R/2+1][R] complex (transposed, shifted rows), gauss [r][r/2+1] -> out [r][r] real."""
    M = max(r // 8, 1)
    rh = r // 2 + 1
    if r == 4:
        # one thread per FFT, DFT-4
        Y = np.zeros((r, rh), complex)
        for kx in range(rh):
            j = np.arange(4)
            cj = (j + r // 2 - 1) & (r - 1)
            s = R // 2 - r // 2 + cj
            vals = spec[kx, s] * gauss[cj, kx]
            Y[:, kx] = dftm(vals[None, :], +1)[0]
        out = np.zeros((r, r))
        for y in range(r // 2):
            Z = np.zeros(r, complex)
            Z[0] = Y[y, 0].real + 1j * Y[y + r // 2, 0].real
            Z[r // 2] = Y[y, r // 2].real + 1j * Y[y + r // 2, r // 2].real
            for kk in range(1, r // 2):
                ya, yb = Y[y, kk], Y[y + r // 2, kk]
                Z[kk] = ya + 1j * yb
                Z[r - kk] = np.conj(ya) + 1j * np.conj(yb)
            o = dftm(Z[None, :], +1)[0]
            out[y], out[y + r // 2] = o.real, o.imag
        return out
    # column stage: FFT kx (0..r/2), thread t holds j = t + M*a
    nf = rh
    tt = np.tile(np.arange(M), nf)
    kx = np.repeat(np.arange(nf), M)
    nthr = nf * M
    v = np.zeros((nthr, 8), complex)
    rin = np.zeros(nthr, int)
    rout = np.zeros(nthr, int)
    for a in range(8):
        j = tt + M * a
        cj = (j + r // 2 - 1) & (r - 1)
        s = R // 2 - r // 2 + cj
        v[:, a] = spec[kx, s] * gauss[cj, kx]
    regs = fft_two_stage(v, tt, M, +1, rin, rout)
    # register q holds y = out_index(t, M, q, 0); with rout = 0: y = t + M*q ; q and q+4 differ by r/2
    Zrow = np.zeros((r // 2, r), complex)
    for i in range(nthr):
        for q in range(4):
            y = out_index(tt[i], M, q, 0)
            y2 = out_index(tt[i], M, q + 4, 0)
            assert y2 == y + r // 2 and y < r // 2
            ya, yb = regs[i, q], regs[i, q + 4]
            kk = kx[i]
            if kk == 0 or kk == r // 2:
                Zrow[y, kk] = ya.real + 1j * yb.real
            else:
                Zrow[y, kk] = ya + 1j * yb
                Zrow[y, r - kk] = np.conj(ya) + 1j * np.conj(yb)
    # row stage: FFT pair y (0..r/2-1), thread t holds k = t + M*a
    nf = r // 2
    tt = np.tile(np.arange(M), nf)
    yy = np.repeat(np.arange(nf), M)
    nthr = nf * M
    v = np.zeros((nthr, 8), complex)
    rin = (yy % 4)
    rout = (yy % min(M, 4)) if M > 1 else np.zeros(nthr, int)
    for a in range(8):
        kcol = tt + M * ((a + rin) & 7)
        v[:, a] = Zrow[yy, kcol]
    regs = fft_two_stage(v, tt, M, +1, rin, rout)
    out = np.zeros((r, r))
    for i in range(nthr):
        for q in range(8):
            xx = out_index(tt[i], M, q, rout[i])
            out[yy[i], xx] = regs[i, q].real
            out[yy[i] + r // 2, xx] = regs[i, q].imag
    return out


def ref_forward(x):
    f = np.fft.rfft2(x, norm='forward')
    f = np.concatenate([f[33:], f[:33]], axis=0)          # shgan.py:315-317
    return f.T                                            # [kx][s]


def ref_inverse(spec_sk, gauss, R, r):
    sp = spec_sk.T                                        # [s][kx]
    crop = sp[R // 2 - r // 2:R // 2 + r // 2, :r // 2 + 1] * gauss
    crop = np.concatenate([crop[r - r // 2 - 1:], crop[:r - r // 2 - 1]], axis=0)
    return np.fft.irfft2(crop, s=(r, r), norm='forward')


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 64))
    got, ref = forward_plane(x), ref_forward(x)
    print('forward max err', np.abs(got - ref).max(), 'ref max', np.abs(ref).max())
    spec = rng.standard_normal((33, 64)) + 1j * rng.standard_normal((33, 64))
    for r in (64, 32, 16, 8, 4):
        g = rng.uniform(0.2, 1.0, (r, r // 2 + 1))
        got, ref = inverse_band(spec, g, 64, r), ref_inverse(spec, g, 64, r)
        print('inverse r', r, 'max err', np.abs(got - ref).max(), 'ref max', np.abs(ref).max())
