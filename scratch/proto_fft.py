"""numpy prototype of the warp-level pruned FFT index math (forward + inverse)."""
import numpy as np

def fwd(frame_win, N):
    """frame_win: the win=N/2 windowed samples a[m]. returns X[k], k=0..N/2 of the centred zero-padded N frame."""
    Nz, Nc, R2 = N//2, N//4, N//64          # R2 = 2R : radix of pass A (pruned: R inputs nonzero)
    R = R2//2
    a = frame_win
    z = a[0::2] + 1j*a[1::2]                 # Nc complex (n < Nz/2), rest zero
    # pass A: lane n2, r<R : z[n2 + 32 r]; radix-2R pruned over r -> Y[n2,k1], k1<2R
    Y = np.zeros((32, R2), complex)
    for n2 in range(32):
        for k1 in range(R2):
            Y[n2,k1] = sum(z[n2+32*r]*np.exp(-2j*np.pi*r*k1/R2) for r in range(R))
            Y[n2,k1] *= np.exp(-2j*np.pi*n2*k1/Nz)
    Z = np.zeros(Nz, complex)
    for k1 in range(R2):
        for k2 in range(32):
            Z[k1+R2*k2] = sum(Y[n2,k1]*np.exp(-2j*np.pi*n2*k2/32) for n2 in range(32))
    # split
    X = np.zeros(Nz+1, complex)
    for k in range(0, Nz//2+1):
        Zk = Z[k]; Zm = np.conj(Z[(Nz-k)%Nz])
        Fe2 = Zk+Zm; Fo2 = Zk-Zm
        w = np.exp(-2j*np.pi*k/N)
        T = (-0.5j*w)*Fo2
        Ak = 0.5*Fe2 + T
        Am = np.conj(0.5*Fe2 - T)
        X[k] = (-1j)**k * Ak
        X[Nz-k] = (-1j)**(Nz-k) * Am
    return X

def inv(X, N):
    """X[k] k=0..N/2 -> centre-half samples of irfft (unwindowed), length N/2."""
    Nz, Nc, R2 = N//2, N//4, N//64
    R = R2//2
    Zp = np.zeros(Nz, complex)
    for k in range(0, Nz//2+1):
        Ak = (1j)**k * X[k]
        Am = (1j)**(Nz-k) * X[Nz-k]
        if k == 0:   # DC / Nyquist imag ignored by irfft
            Ak = Ak.real; Am = Am.real
        # Fe[k] = (A[k] + conj A[Nz-k])/2 ; Fo[k] = (A[k]-conj A[Nz-k])/2 * e^{+2 pi i k/N}; Z' = Fe + i Fo
        Fe = 0.5*(Ak+np.conj(Am)); Fo = 0.5*(Ak-np.conj(Am))*np.exp(2j*np.pi*k/N)
        Zp[k] = Fe + 1j*Fo
        # partner: Z'[Nz-k] = conj(Fe) + i conj(Fo)   (Fe,Fo hermitian)
        if k != 0:
            Zp[(Nz-k)] = np.conj(Fe) + 1j*np.conj(Fo)
    # step B': lane (k1): U[k1][n2] = sum_k2 Z'[k1+R2 k2] w32^{-n2 k2}
    U = np.zeros((R2,32), complex)
    for k1 in range(R2):
        for n2 in range(32):
            U[k1,n2] = sum(Zp[k1+R2*k2]*np.exp(2j*np.pi*n2*k2/32) for k2 in range(32))
            U[k1,n2] *= np.exp(2j*np.pi*n2*k1/Nz)
    z = np.zeros(Nc, complex)
    for n2 in range(32):
        for r in range(R):
            z[n2+32*r] = sum(U[k1,n2]*np.exp(2j*np.pi*r*k1/R2) for k1 in range(R2))
    z /= Nz
    a = np.zeros(Nz)
    a[0::2] = z.real; a[1::2] = z.imag
    return a

for N in (512, 1024):
    rs = np.random.RandomState(0)
    a = rs.randn(N//2)
    fr = np.zeros(N); fr[N//4:3*N//4] = a
    ref = np.fft.rfft(fr)
    X = fwd(a, N)
    print(N, 'fwd err', np.abs(X-ref).max())
    Xr = rs.randn(N//2+1) + 1j*rs.randn(N//2+1)
    refi = np.fft.irfft(Xr)[N//4:3*N//4]
    ai = inv(Xr, N)
    print(N, 'inv err', np.abs(ai-refi).max())
