#!/usr/bin/env python3
"""numpy emulation of the two operand containers of the dense kernel (3xTF32 and 3xFP16 hi/lo
splits, the latter with the per-dimension power-of-two scaling) against float64 on five
magnitude regimes.  Output kept in profiles/r1_split_precision_sim.txt."""
import numpy as np
rng=np.random.default_rng(0)
def tf32_rna(x):
    x=np.asarray(x,np.float32); u=x.view(np.uint32).astype(np.uint64)
    u=(u+0x1000)&0xFFFFE000
    return u.astype(np.uint32).view(np.float32)
def tf32_trunc(x):
    x=np.asarray(x,np.float32); u=x.view(np.uint32)&np.uint32(0xFFFFE000)
    return u.view(np.float32)
def split_tf32(v):
    hi=tf32_rna(v); lo=tf32_trunc((v-hi).astype(np.float32)); return hi.astype(np.float64),lo.astype(np.float64)
def split_f16(v):
    hi=v.astype(np.float16).astype(np.float32); lo=(v-hi).astype(np.float32).astype(np.float16).astype(np.float32)
    return hi.astype(np.float64),lo.astype(np.float64)
def run(D,G,T,xscale,varlo,varhi,meanscale,label):
    means=(meanscale*rng.standard_normal((G,D))).astype(np.float32)
    vars_=rng.uniform(varlo,varhi,(G,D)).astype(np.float32)
    iv=(1/vars_).astype(np.float32); miv=(means*iv).astype(np.float32)
    gc=(-0.5*(D*np.log(2*np.pi)+np.log(vars_).sum(1)+(means**2/vars_).sum(1))).astype(np.float32)
    g=rng.integers(0,G,T); x=(means[g]+np.sqrt(vars_[g])*rng.standard_normal((T,D))*xscale).astype(np.float32)
    A=np.concatenate([x,(x*x).astype(np.float32),np.ones((T,1),np.float32)],1)
    B=np.concatenate([miv,(-0.5*iv).astype(np.float32),gc[:,None]],1)
    truth=A.astype(np.float64)@B.astype(np.float64).T
    fp32=(A[:, :D]@miv.T + gc[None,:]) - 0.5*(A[:,D:2*D]@iv.T)   # fp32 reference-like
    out={}
    ah,al=split_tf32(A); bh,bl=split_tf32(B)
    out['3xTF32']=ah@bh.T+al@bh.T+ah@bl.T
    # fp16 with per-dim power-of-2 scaling from model second moments
    s2=(means.astype(np.float64)**2+vars_).mean(0); k=np.round(0.5*np.log2(s2)).astype(int)
    sc=np.concatenate([2.0**-k,2.0**(-2*k),[1.0]]).astype(np.float32)
    As=(A*sc).astype(np.float32); Bs=(B/sc).astype(np.float32)
    ah,al=split_f16(As); bh,bl=split_f16(Bs)
    out['3xFP16 scaled']=ah@bh.T+al@bh.T+ah@bl.T
    ah,al=split_f16(A); bh,bl=split_f16(B)
    out['3xFP16 unscaled']=ah@bh.T+al@bh.T+ah@bl.T
    print(f"--- {label}: |ll| median {np.median(np.abs(truth)):.1f}, max|A| {np.abs(As).max():.1f} max|B| {np.abs(Bs).max():.1f}")
    print(f"   fp32 (numpy sgemm) max abs err {np.abs(fp32-truth).max():.2e}")
    for n,v in out.items():
        e=np.abs(v-truth); print(f"   {n:16s} max abs err {e.max():.2e}  mean {e.mean():.2e}  inf/nan {np.sum(~np.isfinite(v))}")
run(40,2000,500,1.0,0.5,2.0,3.0,"bench-like synthetic")
run(40,2000,500,1.0,0.01,100.0,30.0,"wide dynamic range (var 0.01..100, means x30)")
run(40,2000,500,3.0,0.5,2.0,3.0,"outlier frames (3 sigma scale)")
run(40,2000,500,1.0,1e-3,1e-2,1.0,"tiny variances (floored region)")
run(39,2000,500,1.0,50,2000,80.0,"MFCC-like large raw magnitudes")
