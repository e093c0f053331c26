"""numpy prototype of the in-warp 512-point real FFT used by csrc/sg_feat.cu (index maps + twiddles)
and of the adjoint of the real-FFT untangle step.  Development aid; not imported by the product."""
import numpy as np

M = 256
W = lambda N, e: np.exp(-2j * np.pi * e / N)


def fft256_lanes(zin):
    """zin[lane][n0] = z[32*n0 + lane].  Returns zout[lane][r] with r = j*4 + k2 holding
    Z[k0 + 8*k1 + 64*k2], lane = k0 + 8*(k1&3), j = k1>>2."""
    # pass A: lane=(n1,n2): lane = n1*4+n2 ; regs n0 -> k0
    A = np.zeros((32, 8), complex)
    for lane in range(32):
        n1 = lane >> 2
        for k0 in range(8):
            A[lane, k0] = sum(zin[lane, n0] * W(8, n0 * k0) for n0 in range(8)) * W(64, n1 * k0)
    # exchange 1: smem[k0*36 + lane]
    sm = np.zeros(8 * 36, complex)
    for lane in range(32):
        for k0 in range(8):
            sm[k0 * 36 + lane] = A[lane, k0]
    Bv = np.zeros((32, 8), complex)
    for lane in range(32):
        k0, n2 = lane >> 2, lane & 3
        v = [sm[k0 * 36 + n1 * 4 + n2] for n1 in range(8)]
        for k1 in range(8):
            Bv[lane, k1] = sum(v[n1] * W(8, n1 * k1) for n1 in range(8)) * W(256, n2 * (k0 + 8 * k1))
    # exchange 2: smem[k1*33 + k0*4 + n2] = smem[k1*33 + lane]
    sm = np.zeros(8 * 33, complex)
    for lane in range(32):
        for k1 in range(8):
            sm[k1 * 33 + lane] = Bv[lane, k1]
    out = np.zeros((32, 8), complex)
    for lane in range(32):
        k0, c = lane & 7, lane >> 3
        for j in range(2):
            k1 = c + 4 * j
            v = [sm[k1 * 33 + k0 * 4 + n2] for n2 in range(4)]
            for k2 in range(4):
                out[lane, j * 4 + k2] = sum(v[n2] * W(4, n2 * k2) for n2 in range(4))
    return out


def out_index(lane, r):
    k0, c = lane & 7, lane >> 3
    j, k2 = r >> 2, r & 3
    return k0 + 8 * (c + 4 * j) + 64 * k2


def test_fft():
    rng = np.random.default_rng(0)
    z = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    zin = np.array([[z[32 * n0 + lane] for n0 in range(8)] for lane in range(32)])
    out = fft256_lanes(zin)
    Z = np.zeros(M, complex)
    for lane in range(32):
        for r in range(8):
            Z[out_index(lane, r)] = out[lane, r]
    assert np.allclose(Z, np.fft.fft(z)), np.abs(Z - np.fft.fft(z)).max()
    print("fft256 lane map OK")


def untangle(Z):
    k = np.arange(M + 1)
    Zk = np.concatenate([Z, Z[:1]])
    Zm = np.conj(Zk[::-1])
    Wk = W(2 * M, k)
    return 0.5 * (Zk + Zm) - 0.5j * Wk * (Zk - Zm)


def untangle_adj(dX):
    k = np.arange(M + 1)
    Wk = W(2 * M, k)
    a, b = 0.5 * (1 - 1j * Wk), 0.5 * (1 + 1j * Wk)
    dZ = np.zeros(M, complex)
    for kp in range(1, M):
        dZ[kp] = np.conj(a[kp]) * dX[kp] + b[M - kp] * np.conj(dX[M - kp])
    dZ[0] = np.conj(a[0]) * dX[0] + np.conj(a[M]) * dX[M] + b[0] * np.conj(dX[0]) + b[M] * np.conj(dX[M])
    return dZ


def test_rfft_and_adjoint():
    rng = np.random.default_rng(1)
    g = rng.standard_normal(2 * M)
    z = g[0::2] + 1j * g[1::2]
    X = untangle(np.fft.fft(z))
    assert np.allclose(X, np.fft.rfft(g))
    # adjoint: L = sum(wr*Re X + wi*Im X)
    wr, wi = rng.standard_normal(M + 1), rng.standard_normal(M + 1)
    dX = wr + 1j * wi
    dZ = untangle_adj(dX)
    dz = np.conj(np.fft.fft(np.conj(dZ)))          # unnormalised inverse via conj trick
    dg = np.zeros(2 * M)
    dg[0::2], dg[1::2] = dz.real, dz.imag
    j = np.arange(2 * M)[:, None]
    bb = np.arange(M + 1)[None, :]
    th = 2 * np.pi * j * bb / (2 * M)
    ref = (wr[None] * np.cos(th) - wi[None] * np.sin(th)).sum(1)
    assert np.allclose(dg, ref), np.abs(dg - ref).max()
    print("rfft untangle + adjoint OK")


if __name__ == "__main__":
    test_fft()
    test_rfft_and_adjoint()
