"""Lane-accurate numpy emulation of csrc/fftcore.cuh (index conventions check)."""
import numpy as np
def brev(x,bits):
    r=0
    for i in range(bits): r|=((x>>i)&1)<<(bits-1-i)
    return r
def ilog2(x): return x.bit_length()-1
def w32(I,inv): 
    return np.exp((2j if inv else -2j)*np.pi*I/32)
def dif(v,LEN,BASE,STRIDE,INV):
    # v: [32 lanes, NV] complex
    if LEN>1:
        H=LEN//2
        for i in range(H):
            a=v[:,BASE+STRIDE*i].copy(); b=v[:,BASE+STRIDE*(i+H)].copy()
            v[:,BASE+STRIDE*i]=a+b
            v[:,BASE+STRIDE*(i+H)]=(a-b)*w32(i*(32//LEN),INV)
        dif(v,H,BASE,STRIDE,INV); dif(v,H,BASE+STRIDE*H,STRIDE,INV)
class Cfg:
    def __init__(s,N):
        s.N=N; s.Nz=N//2; s.R=N//128; s.R2=2*s.R; s.Q=2048//N; s.ZS=s.Nz+(s.R2 if s.R2<16 else 0); s.PI=s.Nz//64
    def posA(s,q,k1): return q*s.R2+(k1&1)*s.R+brev(k1>>1,ilog2(s.R))
def tables(C):
    lane=np.arange(32)
    tw=np.stack([np.exp(-2j*np.pi*k1*lane/C.Nz) for k1 in range(1,C.R2)])  # [(k1-1), lane]
    ws=-0.5j*np.exp(-2j*np.pi*np.arange(C.Nz//2+1)/C.N)
    return tw,ws
def fft_forward(C,v,tw):
    buf=np.zeros(1056,complex); lane=np.arange(32)
    for q in range(C.Q):
        dif(v,C.R,q*C.R2,1,False); dif(v,C.R,q*C.R2+C.R,1,False)
    for k1 in range(C.R2):
        for q in range(C.Q):
            val=v[:,C.posA(q,k1)]*(1 if k1==0 else tw[k1-1])
            buf[lane*33+q*C.R2+k1]=val
    u=np.zeros((32,32),complex)
    for n2 in range(32): u[:,n2]=buf[n2*33+lane]
    dif(u,32,0,1,False)
    buf[:]=np.nan
    zrow=(lane//C.R2)*C.ZS+(lane%C.R2)
    for k2 in range(32): buf[zrow+C.R2*k2]=u[:,brev(k2,5)]
    return buf
def fft_inverse(C,buf,tw):
    lane=np.arange(32); zrow=(lane//C.R2)*C.ZS+(lane%C.R2)
    v=np.zeros((32,32),complex)
    for k2 in range(32): v[:,k2]=buf[zrow+C.R2*k2]
    dif(v,32,0,1,True)
    b2=np.zeros(1056,complex)
    for n2 in range(32): b2[n2*33+lane]=v[:,brev(n2,5)]
    for j in range(32): v[:,j]=b2[lane*33+j]
    for k1 in range(1,C.R2):
        for q in range(C.Q): v[:,q*C.R2+k1]*=np.conj(tw[k1-1])
    out=np.zeros((32,32),complex)
    for q in range(C.Q):
        dif(v,C.R,q*C.R2,2,True); dif(v,C.R,q*C.R2+1,2,True)
        for r in range(C.R):
            p=q*C.R2+2*brev(r,ilog2(C.R))
            out[:,q*C.R+r]=v[:,p]+v[:,p+1]*w32(r*(32//C.R2),True)
    return out
def rot_fwd(a,j): return a*((-1j)**(j&3))
def rot_inv(a,j): return a*((1j)**(j&3))
def split_fwd(Zk,Zr,wsk):
    fe=Zk+np.conj(Zr); fo=Zk-np.conj(Zr); t=wsk*fo
    return 0.5*fe+t, np.conj(0.5*fe-t)
def split_inv(Ak,Am,wsk):
    fe=Ak+np.conj(Am); g=Ak-np.conj(Am); h=g*np.conj(wsk)
    return fe+2*h, np.conj(fe-2*h)

for N in (512,1024,2048):
    C=Cfg(N); tw,ws=tables(C); rs=np.random.RandomState(N)
    a=rs.randn(C.Q,N//2)   # windowed samples per frame
    lane=np.arange(32)
    v=np.zeros((32,32),complex)
    for q in range(C.Q):
        for r in range(C.R):
            m=2*(lane+32*r); z=a[q,m]+1j*a[q,m+1]
            v[:,q*C.R2+r]=z; v[:,q*C.R2+C.R+r]=z*w32(r*(32//C.R2),False)
    buf=fft_forward(C,v,tw)
    X=np.zeros((C.Q,C.Nz+1),complex)
    for q in range(C.Q):
        for i in range(C.PI):
            k=lane+32*i
            Ak,Am=split_fwd(buf[q*C.ZS+k],buf[q*C.ZS+((C.Nz-k)&(C.Nz-1))],ws[k])
            X[q,k]=rot_fwd(Ak,lane); X[q,C.Nz-k]=[rot_fwd(Am[l],(C.Nz-k[l])) for l in range(32)]
        k=C.Nz//2
        Ak,_=split_fwd(buf[q*C.ZS+k],buf[q*C.ZS+k],ws[k]); X[q,k]=rot_fwd(Ak,k)
    fr=np.zeros((C.Q,N)); fr[:,N//4:3*N//4]=a
    ref=np.fft.rfft(fr,axis=1)
    print(N,'fwd',np.abs(X-ref).max())
    # inverse
    Xr=rs.randn(C.Q,C.Nz+1)+1j*rs.randn(C.Q,C.Nz+1)
    buf=np.full(1056,np.nan,complex)
    for q in range(C.Q):
        for i in range(C.PI):
            k=lane+32*i
            Ak=np.array([rot_inv(Xr[q,k[l]],k[l]) for l in range(32)])
            Am=np.array([rot_inv(Xr[q,C.Nz-k[l]],C.Nz-k[l]) for l in range(32)])
            if i==0: Ak[0]=Ak[0].real; Am[0]=Am[0].real
            Zk,Zr=split_inv(Ak,Am,ws[k])
            buf[q*C.ZS+((C.Nz-k)&(C.Nz-1))]=Zr   # write partner first; for k=0 slot 0 overwritten by Zk
            buf[q*C.ZS+k]=Zk
        k=C.Nz//2; A=rot_inv(Xr[q,k],k); Zk,_=split_inv(A,A,ws[k]); buf[q*C.ZS+k]=Zk
    out=fft_inverse(C,buf,tw)/N
    rec=np.zeros((C.Q,N//2))
    for q in range(C.Q):
        for r in range(C.R):
            m=2*(lane+32*r); rec[q,m]=out[:,q*C.R+r].real; rec[q,m+1]=out[:,q*C.R+r].imag
    refi=np.fft.irfft(Xr,axis=1)[:,N//4:3*N//4]
    print(N,'inv',np.abs(rec-refi).max())
