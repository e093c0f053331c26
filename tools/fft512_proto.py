"""numpy prototype of the in-warp 512-point complex FFT (1024-point real FFT of AudioNet's STFT,
Preprocessor.py:100-105) used by csrc/sg_audionet.cu: index maps, twiddles, smem paddings."""
import itertools
import numpy as np

W = lambda N, e: np.exp(-2j * np.pi * e / N)


def conflicts(addr_fn):
    """max bank multiplicity over the 32 lanes for every fixed (h, reg)."""
    worst = 1
    for h in range(2):
        for r in range(8):
            banks = {}
            for lane in range(32):
                b = addr_fn(lane, h, r) % 32
                banks[b] = banks.get(b, 0) + 1
            worst = max(worst, max(banks.values()))
    return worst


def search():
    best = None
    for S1, P1 in itertools.product(range(64, 100), range(8, 12)):
        # exchange 1: index (k0,n1,n2): addr = k0*S1 + n1*P1 + n2
        w = conflicts(lambda lane, h, k0: k0 * S1 + (((lane + 32 * h) >> 3)) * P1 + ((lane + 32 * h) & 7))
        r = conflicts(lambda lane, h, n1: ((lane >> 3) + 4 * h) * S1 + n1 * P1 + (lane & 7))
        if w == 1 and r == 1:
            best = (S1, P1)
            break
    print("exchange1 (S1,P1):", best)
    best2, bw = None, 99
    for S2, P2 in itertools.product(range(64, 140), range(8, 17)):
        # exchange 2: index (k0,k1,n2): addr = k1*S2 + k0*P2 + n2 ; writer (k0=(lane>>3)+4h, n2=lane&7), reg k1
        w = conflicts(lambda lane, h, k1: k1 * S2 + ((lane >> 3) + 4 * h) * P2 + (lane & 7))
        # reader: k0 = lane&7, k1 = (lane>>3)+4h, reg n2
        r = conflicts(lambda lane, h, n2: ((lane >> 3) + 4 * h) * S2 + (lane & 7) * P2 + n2)
        if w + r < bw and S2 >= 8 * P2:
            best2, bw = (S2, P2), w + r
            print("  cand", S2, P2, "write-way", w, "read-way", r)
        if w == 1 and r == 1:
            break
    print("exchange2 (S2,P2):", best2)
    return best, best2


def fft512_lanes(z, S1, P1, S2, P2):
    """z[512] complex -> Z[512], emulating the lane/register/smem choreography."""
    # pass A: thread (lane,h): e = lane+32h, n1 = e>>3, n2 = e&7 ; regs n0
    A = np.zeros((32, 2, 8), complex)
    for lane in range(32):
        for h in range(2):
            e = lane + 32 * h
            n1 = e >> 3
            v = np.array([z[64 * n0 + e] for n0 in range(8)])
            for k0 in range(8):
                A[lane, h, k0] = sum(v[n0] * W(8, n0 * k0) for n0 in range(8)) * W(64, n1 * k0)
    sm = np.zeros(8 * S1 + 64, complex)
    for lane in range(32):
        for h in range(2):
            e = lane + 32 * h
            for k0 in range(8):
                sm[k0 * S1 + (e >> 3) * P1 + (e & 7)] = A[lane, h, k0]
    Bv = np.zeros((32, 2, 8), complex)
    for lane in range(32):
        for h in range(2):
            k0, n2 = (lane >> 3) + 4 * h, lane & 7
            v = [sm[k0 * S1 + n1 * P1 + n2] for n1 in range(8)]
            for k1 in range(8):
                Bv[lane, h, k1] = sum(v[n1] * W(8, n1 * k1) for n1 in range(8)) * W(512, n2 * (k0 + 8 * k1))
    sm = np.zeros(8 * S2 + 64, complex)
    for lane in range(32):
        for h in range(2):
            k0, n2 = (lane >> 3) + 4 * h, lane & 7
            for k1 in range(8):
                sm[k1 * S2 + k0 * P2 + n2] = Bv[lane, h, k1]
    Z = np.zeros(512, complex)
    for lane in range(32):
        for h in range(2):
            k0, k1 = lane & 7, (lane >> 3) + 4 * h
            v = [sm[k1 * S2 + k0 * P2 + n2] for n2 in range(8)]
            for k2 in range(8):
                Z[k0 + 8 * k1 + 64 * k2] = sum(v[n2] * W(8, n2 * k2) for n2 in range(8))
    return Z


if __name__ == "__main__":
    (S1, P1), (S2, P2) = search()
    rng = np.random.default_rng(0)
    z = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    Z = fft512_lanes(z, S1, P1, S2, P2)
    print("fft512 err", np.abs(Z - np.fft.fft(z)).max())
    # output scatter: thread (lane,h) reg k2 writes Z[k0 + 8*k1 + 64*k2] with k0=lane&7, k1=(lane>>3)+4h -> index lane+32h+64k2
    print("out index of (lane=5,h=1,k2=3):", (5 & 7) + 8 * ((5 >> 3) + 4) + 64 * 3, "==", 5 + 32 + 192)
