"""Numpy model of the EXPERIMENTAL warp-level band transform of d4c_body_kernel<12, true> (wb_d4c.cu): the 4096-point
transform of a 513-sample slice as eight 512-point transforms (residue r = warp), lane / staging / output index
mapping exactly as in the kernel.  Prints the relative error against numpy's FFT (6e-15)."""
import numpy as np
rng=np.random.default_rng(0)
N=4096
z=np.zeros(N,complex); z[:513]=rng.normal(size=513)+1j*rng.normal(size=513)
Z=np.fft.ifft(z)*N   # forward with e^{+i}: sum z[n] e^{+2 pi i n k / N}
T=np.exp(2j*np.pi*np.arange(2*N)/(2*N))   # table of 8192 entries
def dft16(a):  # a[p] = sum_t a[t] W16^{tp}, W16 = e^{+2 pi i/16}
    t=np.arange(16)
    return np.array([np.sum(a*np.exp(2j*np.pi*t*p/16)) for p in range(16)])
out=np.zeros(N,complex)
for r in range(8):          # warp r
    # modulated input, lane l holds n = l + 32 t
    u=np.zeros(512,complex)
    n=np.arange(512)
    u[:]=z[:512]*T[(2*n*r)%8192]
    u[0]+=z[512]*T[(2*512*r)%8192]      # wrap term: W_4096^{512 r}
    stage=np.zeros(16*34,complex)
    for l in range(32):     # pass A
        a=dft16(u[l+32*np.arange(16)])
        w=T[(16*l)%8192]    # W_512^{l}
        a=a*w**np.arange(16)
        for p in range(16): stage[p*34+l]=a[p]
    res={}
    for lp in range(32):    # pass B: lane lp = 2p + h
        p,h=lp>>1,lp&1
        a=dft16(stage[p*34+h+2*np.arange(16)])
        w=T[(256*h)%8192]   # W_32^{h}
        a=a*w**np.arange(16)
        res[lp]=a
    for lp in range(32):    # radix-2 across lane pairs (shuffle xor 1)
        p,h=lp>>1,lp&1
        other=res[lp^1]
        mine=res[lp]
        val = (mine+other) if h==0 else (other-mine)   # h=0: y0+y1 ; h=1: y0-y1 (other = y0)
        for p2 in range(16):
            q=p+16*p2+256*h
            out[8*q+r]=val[p2]
print(np.max(np.abs(out-Z))/np.max(np.abs(Z)))
