"""Index maps of the half-warp-per-frame FFT of sg_feat.cu (V2 kernels) checked against numpy: the radix-4 x 4 fft16 with its output
permutation, 256 = 16 x 16 over 16 lanes with one exchange, and the real-FFT untangle with the mirror bin in the partner lane."""
import numpy as np
rng=np.random.default_rng(0)
def fft4(a0,a1,a2,a3):
    t0,t1,t2,t3=a0+a2,a0-a2,a1+a3,a1-a3
    return t0+t2, t1-1j*t3, t0-t2, t1+1j*t3
def fft16(v):
    v=list(v)
    for n0 in range(4):
        v[n0],v[n0+4],v[n0+8],v[n0+12]=fft4(v[n0],v[n0+4],v[n0+8],v[n0+12])
    for n0 in range(1,4):
        for k1 in range(1,4):
            v[n0+4*k1]*=np.exp(-2j*np.pi*n0*k1/16)
    for k1 in range(4):
        v[4*k1],v[4*k1+1],v[4*k1+2],v[4*k1+3]=fft4(v[4*k1],v[4*k1+1],v[4*k1+2],v[4*k1+3])
    t=[None]*16
    for k1 in range(4):
        for k2 in range(4): t[k1+4*k2]=v[4*k1+k2]
    return t
x=rng.normal(size=16)+1j*rng.normal(size=16)
print('fft16 err',np.abs(np.array(fft16(x))-np.fft.fft(x)).max())
# 256 = 16x16 across 16 lanes
x=rng.normal(size=256)+1j*rng.normal(size=256)
z=[[x[16*a+l] for a in range(16)] for l in range(16)]      # lane l slot a
z=[fft16(z[l]) for l in range(16)]                          # slot k1
for l in range(16):
    for k1 in range(16): z[l][k1]*=np.exp(-2j*np.pi*l*k1/256)
ex=np.zeros((16,16),complex)
for l in range(16):
    for k1 in range(16): ex[k1,l]=z[l][k1]
z=[[ex[l,b] for b in range(16)] for l in range(16)]        # lane k1=l slot b
z=[fft16(z[l]) for l in range(16)]                          # lane k1 slot k2 -> k = k1+16k2
X=np.fft.fft(x)
err=max(abs(z[l][i]-X[l+16*i]) for l in range(16) for i in range(16))
print('fft256 err',err)
# real FFT untangle with mirror via partner lane
g=rng.normal(size=512); g[400:]=0
zz=g[0::2]+1j*g[1::2]
Z=np.fft.fft(zz)
Zl=[[Z[l+16*i] for i in range(16)] for l in range(16)]
G=np.fft.fft(g)
err=0
for l in range(16):
    for i in range(16):
        k=l+16*i
        if l==0: zm=np.conj(Zl[0][(16-i)&15])
        else: zm=np.conj(Zl[(16-l)&15][15-i])
        c,s=np.cos(2*np.pi*k/512),np.sin(2*np.pi*k/512)
        a=0.5*(1-s)-0.5j*c; b=0.5*(1+s)+0.5j*c
        err=max(err,abs(a*Zl[l][i]+b*zm-G[k]))
print('untangle err',err)
