#!/usr/bin/env python
"""Study (round 2): what numpy's float64 exp really is on an AVX-512 host, and how to reproduce it.

The reference's ESACF peak interpolation evaluates its Gaussian model with numpy.exp inside
scipy.optimize.curve_fit (peakutils.gaussian, esacf.py:60-62).  On this image numpy dispatches
float64 exp to Intel SVML's __svml_exp8_ha (6 call sites in _multiarray_umath.so, none to glibc's
exp), whose main path is 20 AVX-512 instructions: N = floor16(x log2 e) via one fused multiply-add
in round-toward-zero with a shifter constant, r = x - N ln2 (hi / lo), a degree-6 polynomial in r
evaluated in Estrin form, two 16-entry tables 2^(j/16) (high and low part) and a scalef.  The
constants below were read out of the shared object's `__svml_dexp_ha_data_internal_avx512` block.
This emulation (exact rational arithmetic for the fused operations) reproduces numpy.exp BIT FOR
BIT on 26 000 arguments in (-700, 50); glibc's exp and a correctly rounded exp differ from it on
4.6 % of arguments.  Consequence: the chaotic ("runaway") Levenberg-Marquardt fits of
the reference depend on the CPU's SIMD dispatch, and a bit-exact device replica would additionally
need glibc's pow(): the model's `dev ** 2` is a numpy SCALAR power = libm pow(dev, 2.0), which
differs from dev * dev on 0.084 % of arguments (measured, 300 000 values).
Run: python scripts/studies/svml_exp_emulation.py   (needs an AVX-512 host to compare against)."""
import numpy as np, math, struct
from fractions import Fraction as Fr
H=float.fromhex
TH=[H(v) for v in ['0x1.0000000000000p+0','0x1.0b5586cf9890fp+0','0x1.172b83c7d517bp+0','0x1.2387a6e756238p+0','0x1.306fe0a31b715p+0','0x1.3dea64c123422p+0','0x1.4bfdad5362a27p+0','0x1.5ab07dd485429p+0','0x1.6a09e667f3bcdp+0','0x1.7a11473eb0187p+0','0x1.8ace5422aa0dbp+0','0x1.9c49182a3f090p+0','0x1.ae89f995ad3adp+0','0x1.c199bdd85529cp+0','0x1.d5818dcfba487p+0','0x1.ea4afa2a490dap+0']]
TL=[H(v) for v in ['0x0.0p+0','0x1.79aa65d837b6dp-54','-0x1.01b15eaa59348p-55','0x1.68efde3a8a894p-54','0x1.34d754db0abb6p-55','0x1.59f48a72a4c6dp-55','0x1.690cebb7aafb0p-56','0x1.063e1e21c5409p-54','-0x1.3b3efbf5e2228p-54','-0x1.b32dcb94da51dp-56','0x1.db72fc1f0eab4p-55','0x1.1affc2b91ce27p-56','0x1.c1a7792cb3387p-55','0x1.36eae30af0cb3p-56','0x1.4a385a63d07a7p-56','-0x1.ff7128fd391f0p-55']]
L2E=H('0x1.71547652b82fep+0'); SH=H('0x1.8000000003ff0p+48'); L2H=H('0x1.62e42fefa39efp-1'); L2L=H('0x1.abc9e3b39803fp-56')
c6=H('0x1.7411836940c04p-10'); c12=H('0x1.1101cbbc265c0p-7'); c7=H('0x1.55557242d68fep-5'); c9=H('0x1.5555553939732p-3'); c8=H('0x1.000000000d008p-1'); c11=H('0x1.fffffffffff70p-1')
def fma(a,b,c): return float(Fr(a)*Fr(b)+Fr(c))
def fma_rz(a,b,c):
    e=Fr(a)*Fr(b)+Fr(c)
    f=float(e)  # nearest
    if Fr(f)==e: return f
    # toward zero
    if abs(Fr(f))>abs(e): f=math.nextafter(f,0.0)
    return f
def svml_exp(x):
    z=fma_rz(x,L2E,SH)
    bits=struct.unpack('<Q',struct.pack('<d',z))[0]
    j=bits&15
    N=z-SH
    r=fma(-N,L2H,x)
    r=fma(-N,L2L,r)
    r2=r*r
    p12=fma(c6,r,c12); p9=fma(c7,r,c9); p11=fma(c8,r,c11)
    p12=fma(p12,r2,p9); p12=fma(p12,r2,p11)
    t=fma(r,p12,TL[j])
    res=fma(TH[j],t,TH[j])
    return math.ldexp(res, math.floor(N))
if __name__=='__main__':
    rng=np.random.default_rng(2)
    xs=np.concatenate([-rng.uniform(0,700,20000), rng.uniform(0,50,2000), -rng.uniform(0,1e-3,1000), -10.0**rng.uniform(-12,2.8,3000)])
    ref=np.exp(xs)
    got=np.array([svml_exp(float(v)) for v in xs])
    print("mismatch frac",np.mean(got!=ref), "n",len(xs))
    bad=np.nonzero(got!=ref)[0][:5]
    for i in bad: print(xs[i], got[i].hex(), ref[i].hex())
