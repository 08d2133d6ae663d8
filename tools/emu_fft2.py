#!/usr/bin/env python3
"""numpy emulation of the packed engine's index algebra (fft2.cuh / feat2.cuh): pruned DIT pass A, twiddle,
32x32 transpose, DIT pass B, Hermitian split with the lane-partner exchange.  Checks against numpy.fft.rfft."""
import numpy as np

def brev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

def bfly(a, b, I, inv):
    th = 2 * np.pi * I / 32
    c, s = np.cos(th), (-np.sin(th) if inv else np.sin(th))
    w = c - 1j * s
    if I == 0:
        return a + b, a - b
    if abs(c) >= abs(s):
        t = s / c
        tr = b.real + t * b.imag; ti = b.imag - t * b.real
        return (a.real + c * tr) + 1j * (a.imag + c * ti), (a.real - c * tr) + 1j * (a.imag - c * ti)
    k = c / s
    tr = k * b.real + b.imag; tn = b.real - k * b.imag
    return (a.real + s * tr) + 1j * (a.imag - s * tn), (a.real - s * tr) + 1j * (a.imag + s * tn)

def dit(v, L, base, inv, min_m):
    if L > 1:
        H = L // 2
        dit(v, H, base, inv, min_m); dit(v, H, base + H, inv, min_m)
        if L >= min_m:
            for j in range(H):
                v[base + j], v[base + j + H] = bfly(v[base + j], v[base + j + H], j * (32 // L), inv)

def run(N, hop, seed=0):
    rng = np.random.RandomState(seed)
    Nz, R, R2, P = N // 2, N // 128, N // 64, 2048 // N
    logR2 = int(np.log2(R2))
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N // 2) / (N // 2))
    frames = rng.randn(P, N // 2)                      # P "pairs" (one frame each is enough for index checks)
    a = frames * win * 0.5
    z = a[:, 0::2] + 1j * a[:, 1::2]                   # [P, Nz/2]
    V = np.zeros((32, 32), complex)                    # [lane][reg]
    for lane in range(32):
        v = [0j] * 32
        for p in range(P):
            for r in range(R):
                idx = p * R2 + brev(r, logR2)
                v[idx] = z[p, lane + 32 * r]; v[idx + 1] = v[idx]
            dit(v, R2, p * R2, False, 4)
        for p in range(P):
            for k1 in range(R2):
                v[p * R2 + k1] *= np.exp(-2j * np.pi * ((k1 * lane) % Nz) / Nz)
        V[lane] = v
    X = V.copy()                                       # X[row=lane][col]
    Z = np.zeros((32, 32), complex)
    for lane in range(32):
        v = [0j] * 32
        for n1 in range(32):
            v[brev(n1, 5)] = X[n1][lane]
        dit(v, 32, 0, False, 2)
        Z[lane] = v
    # check Z against direct FFT
    for lane in range(32):
        p, k1 = lane // R2, lane % R2
        zz = np.zeros(Nz, complex); zz[:Nz // 2] = z[p]
        ref = np.fft.fft(zz)
        got = Z[lane]
        assert np.allclose(got, ref[k1 + R2 * np.arange(32)], atol=1e-9), ("Z", N, lane)
    # split
    out = np.zeros((P, Nz + 1), complex)
    for lane in range(32):
        p, k1 = lane // R2, lane % R2
        partner = (lane & ~(R2 - 1)) | ((R2 - k1) & (R2 - 1))
        col0p = (partner % R2) == 0
        for s in range(17):
            if s < 16:
                send = Z[partner][(32 - s) & 31] if col0p else Z[partner][31 - s]   # what the partner lane sends
                Zk, Zr = Z[lane][s], send
            else:
                Zk = Zr = Z[lane][16]
            k = k1 + R2 * s
            ph = 2 * np.pi * k / N
            fe = Zk + np.conj(Zr); fo = Zk - np.conj(Zr)
            fr, fi = fo.real, fo.imag
            if s < 8:
                t, c = np.tan(ph), np.cos(ph)
                tr = fi - t * fr; tn = fr + t * fi
                ak = (fe.real + c * tr) + 1j * (fe.imag - c * tn); am = (fe.real - c * tr) + 1j * (fe.imag + c * tn)
            else:
                ct, sn = np.cos(ph) / np.sin(ph), np.sin(ph)
                nr = fr - ct * fi; tn = fi + ct * fr
                ak = (fe.real - sn * nr) + 1j * (fe.imag - sn * tn); am = (fe.real + sn * nr) + 1j * (fe.imag + sn * tn)
            if s < 16:
                out[p, k] = ak; out[p, Nz - k] = np.conj(am)
            elif k1 == 0:
                out[p, Nz // 2] = ak
    for p in range(P):
        full = np.zeros(N); full[N // 4: N // 4 + N // 2] = frames[p] * win
        ref = np.fft.rfft(full)
        k = np.arange(Nz + 1)
        got = out[p] * ((-1j) ** k)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-12, (N, p, err)
    return True

for N, hop in ((2048, 256), (1024, 120), (512, 60)):
    run(N, hop)
    print("ok", N)
