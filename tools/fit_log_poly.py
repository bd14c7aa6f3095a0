import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pn
lo, hi = np.sqrt(0.5)-1, np.sqrt(2)-1
# P(x) = (log1p(x) - x + x^2/2)/x^3 ; fit degree-8 on Chebyshev nodes (near-minimax), then Remez-like refine by weighted LS
def target(x):
    x = np.asarray(x, dtype=np.longdouble)
    return ((np.log1p(x) - x + x*x/2)/(x**3)).astype(np.float64)
N=4000
k=np.arange(N)
t=np.cos(np.pi*(k+0.5)/N)
x=(lo+hi)/2+(hi-lo)/2*t
x=x[np.abs(x)>1e-3]
y=target(x)
deg=8
# weight: error in log is x^3*dP, relative to |log| ~ |x| -> weight x^2
w=np.abs(x)**2
for it in range(30):
    V=np.vander(x,deg+1,increasing=True)
    c,res,rk,sv=np.linalg.lstsq(V*w[:,None],y*w,rcond=None)
    err=(V@c-y)*np.abs(x)**2
    # Lawson iteration towards minimax
    w=w*np.sqrt(np.abs(err)/np.abs(err).max()+1e-3)
c32=c.astype(np.float32)
print("coeffs (c0..c8):", [repr(float(v)) for v in c32])
# accuracy test in float32 emulation using float64 ops w/ rounding to f32 after each fma (fma exact then round)
def f32(v): return np.float32(v)
def fma(a,b,c): return np.float32(np.float64(a)*np.float64(b)+np.float64(c))   # double product of two f32 is exact; sum rounding double then float: double rounding negligible for test
def mylog(xs):
    xs=np.asarray(xs,dtype=np.float32)
    ix=xs.view(np.uint32).astype(np.uint64)
    ix=(ix+np.uint64(0x3f800000-0x3f3504f3))&np.uint64(0xffffffff)
    k=(ix>>np.uint64(23)).astype(np.int64)-0x7f
    ix=(ix&np.uint64(0x007fffff))+np.uint64(0x3f3504f3)
    m=ix.astype(np.uint32).view(np.float32)
    f=(m-np.float32(1)).astype(np.float32)
    p=np.full_like(f,c32[8])
    for j in range(7,-1,-1): p=fma(p,f,c32[j])
    z=(f*f).astype(np.float32)
    dk=k.astype(np.float32)
    y=((f*z).astype(np.float32)*p).astype(np.float32)
    y=fma(dk,np.float32(9.0580006145e-06),y)
    y=fma(np.float32(-0.5),z,y)
    r=(f+y).astype(np.float32)
    return fma(dk,np.float32(6.9313812256e-01),r)
rng=np.random.default_rng(1)
u=((rng.integers(0,2**24,2000000)+0.5)*2.0**-24).astype(np.float32)
e=np.exp(rng.uniform(np.log(1e-3),np.log(30),1000000)).astype(np.float32)
for name,xs in (("u",u),("E",e)):
    ref=np.log(xs.astype(np.float64))
    got=mylog(xs).astype(np.float64)
    abs_err=np.abs(got-ref)
    ulp=np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    sel=np.abs(ref)>1e-6
    print(name,"max abs err",abs_err.max(),"max ulp err",(abs_err[sel]/ulp[sel]).max(), "max rel", (abs_err[sel]/np.abs(ref[sel])).max())
