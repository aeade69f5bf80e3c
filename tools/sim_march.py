"""Warp-level simulation of lol_render's marches on scene4.lol (numpy, float32, no GPU): the numbers DESIGN.md
quotes for what was NOT built, and why.

    python tools/sim_march.py 1920 1080 shadow   # lane efficiency of the primary / shadow loops per 8x4 tile,
                                                  # both lights in one loop (no gain), how often the shadow
                                                  # division changes res (one step in seven), deferring long rays
                                                  # (14-15 % fewer warp-iterations with ideal repacking)
    python tools/sim_march.py 960 540 far        # Lipschitz caching of the blob's distance between steps
                                                  # (the box test stays in 93 % of warp-iterations: -2.6 % modelled)

It restates scene4's distance field and the kernel's march rules (exact skips on) in plain numpy: the evaluation
counts it prints (21.9 primary, 31.8 shadow evaluations per pixel) equal the kernel's counters to three digits, which
is the check that it marches like the kernel.  Statistics only -- nothing here rounds like the reference.
"""
import sys

MODE = sys.argv[3] if len(sys.argv) > 3 else "shadow"

if MODE == 'shadow':
    import numpy as np, sys
    W,H = int(sys.argv[1]), int(sys.argv[2])
    f32=np.float32
    # scene4
    S = np.array([[0,1,-6,1],[-1,.5,-3,3],[-3,4.5,-3,.5],[2,2,-10,2],[6,2,-10,5]],f32)
    k=f32(3)
    def smin(a,b):
        h=np.clip(f32(.5)+f32(.5)*(b-a)/k,0,1).astype(f32); return (b+(a-b)*h)-k*h*(1-h)
    def sdf(p):
        t=[np.sqrt(((p-S[i,:3])**2).sum(-1))-S[i,3] for i in range(5)]
        blob=smin(smin(t[0],t[1]),smin(t[2],smin(t[3],t[4])))
        plane=p[...,1]+1
        return np.minimum(blob,plane)
    cam=np.array([-2,6,3],f32); d=np.array([.3,-.7,-1],f32); fov=150
    dirn=d/np.linalg.norm(d); hh=np.arctan(np.radians(fov)/2); ww=W/H*hh
    right=np.cross(dirn,[0,1,0]); right/=np.linalg.norm(right); up=np.cross(right,dirn)
    xs=(np.arange(W)+.5)/W*2-1; ys=1-(np.arange(H)+.5)/H*2
    vx,vy=np.meshgrid(xs,ys)
    rd=right*(vx*ww)[...,None]+up*(vy*hh)[...,None]+dirn; rd/=np.linalg.norm(rd,axis=-1,keepdims=True); rd=rd.astype(f32)
    t=np.zeros((H,W),f32); act=np.ones((H,W),bool); npri=np.zeros((H,W),int)
    for i in range(256):
        dd=sdf(cam+rd*t[...,None]); dd=np.where(act,dd,0); t=t+dd; npri+=act
        act=act&~((dd<1e-3)|(t>100))
        if not act.any(): break
    hit=t<100
    p=cam+rd*t[...,None]
    h=(t/100)[...,None]
    ks=np.array([[1,-1,-1],[-1,-1,1],[-1,1,-1],[1,1,1]],f32)
    n=sum(kk*sdf(p+kk*h)[...,None] for kk in ks); n/=np.linalg.norm(n,axis=-1,keepdims=True)
    L=np.array([[-2,10,-1],[-7,2,-5]],f32)
    steps=[]
    for li in range(2):
        l=L[li]-p; ld=np.linalg.norm(l,axis=-1); l/=ld[...,None]
        ndl=(n*l).sum(-1)
        act=hit&(ndl>0)
        so=p+l; st=np.zeros((H,W),f32); res=np.ones((H,W),f32); ns=np.zeros((H,W),int)
        for i in range(128):
            dd=sdf(so+l*st[...,None])
            with np.errstate(all='ignore'): q=50*dd/st
            res=np.where(act,np.fmin(res,q),res); st=np.where(act,st+dd,st); ns+=act
            act=act&~((res<-1)|(st>ld)|(res<=0))
            if not act.any(): break
        steps.append(ns)
    s0,s1=steps
    print("hit frac",hit.mean(),"shadow evals/pixel",(s0+s1).mean(), "primary", npri.mean())
    def tiles(a,th=4,tw=8):
        Hc,Wc=a.shape[0]//th*th,a.shape[1]//tw*tw
        return a[:Hc,:Wc].reshape(Hc//th,th,Wc//tw,tw).transpose(0,2,1,3).reshape(-1,th*tw)
    T0,T1,TP=tiles(s0),tiles(s1),tiles(npri)
    cur=T0.max(1)+T1.max(1); comb=(T0+T1).max(1); ideal=(T0+T1).sum(1)/32
    print("warp shadow iterations: separate loops",cur.sum()," combined",comb.sum()," ideal(perfect packing)",ideal.sum())
    print("efficiency separate",ideal.sum()/cur.sum()," combined",ideal.sum()/comb.sum())
    print("primary eff",TP.sum()/32/TP.max(1).sum())
    tot_cur=TP.max(1).sum()+cur.sum(); tot_new=TP.max(1).sum()+comb.sum()
    print("march iterations total: now",tot_cur," combined",tot_new," ratio",tot_new/tot_cur)

    # how often does a warp need the shadow division at all?  (update needed iff q < res)
    tot=0; need_it=0; lane_need=0; lane_tot=0
    for li in range(2):
        l=L[li]-p; ld=np.linalg.norm(l,axis=-1); l/=ld[...,None]
        ndl=(n*l).sum(-1)
        act=hit&(ndl>0)
        so=p+l; st=np.zeros((H,W),f32); res=np.ones((H,W),f32)
        for i in range(128):
            dd=sdf(so+l*st[...,None])
            with np.errstate(all='ignore'): q=50*dd/st
            need=act&(q<res)
            ta=tiles(act).any(1); tn=tiles(need).any(1)
            tot+=ta.sum(); need_it+=tn.sum(); lane_need+=need.sum(); lane_tot+=act.sum()
            res=np.where(act,np.fmin(res,q),res); st=np.where(act,st+dd,st)
            act=act&~((res<-1)|(st>ld)|(res<=0))
            if not act.any(): break
    print("warp shadow iterations",tot,"of which some lane updates res",need_it, need_it/tot, " lane-level", lane_need/lane_tot)

    print("--- deferral of long rays (cap K per loop, remainder repacked densely) ---")
    base = TP.max(1).sum() + T0.max(1).sum() + T1.max(1).sum()
    for K in (8, 12, 16, 24, 32, 48, 64):
        tot = 0
        for T in (TP, T0, T1):
            tot += np.minimum(T.max(1), K).sum() + np.maximum(T - K, 0).sum() / 32.0
        print(K, tot / base)
    for Kp, Ks in ((32, 16), (48, 24), (64, 32)):
        print("  rays beyond the cap: primary", float((TP > Kp).mean()), "shadow", float(((T0 > Ks).sum() + (T1 > Ks).sum()) / max(1, (T0 > 0).sum() + (T1 > 0).sum())),
              " warps with at least one such ray: primary", float((TP.max(1) > Kp).mean()), "shadow", float(((T0.max(1) > Ks).sum() + (T1.max(1) > Ks).sum()) / max(1, (T0.max(1) > 0).sum() + (T1.max(1) > 0).sum())))
        tot = np.minimum(TP.max(1), Kp).sum() + np.maximum(TP - Kp, 0).sum() / 32.0
        for T in (T0, T1):
            tot += np.minimum(T.max(1), Ks).sum() + np.maximum(T - Ks, 0).sum() / 32.0
        print("primary cap", Kp, "shadow cap", Ks, tot / base)

if MODE == 'far':
    import numpy as np, sys
    W,H = int(sys.argv[1]), int(sys.argv[2])
    f32=np.float32
    S = np.array([[0,1,-6,1],[-1,.5,-3,3],[-3,4.5,-3,.5],[2,2,-10,2],[6,2,-10,5]],f32)
    k=f32(3)
    BC=np.array([3.5,2,-7.5],f32); BH=np.array([7.54375887,5.03850746,7.54375887],f32); M1=f32(2.2545011*1.004)
    def smin(a,b):
        h=np.clip(f32(.5)+f32(.5)*(b-a)/k,0,1).astype(f32); return (b+(a-b)*h)-k*h*(1-h)
    def blob(p):
        t=[np.sqrt(((p-S[i,:3])**2).sum(-1))-S[i,3] for i in range(5)]
        return smin(smin(t[0],t[1]),smin(t[2],smin(t[3],t[4])))
    def boxskip(p,best):
        q=np.maximum(np.abs(p-BC)-BH,0); u=best*f32(1.004)+M1
        return (u>0)&((q*q).sum(-1)>u*u)
    def tiles(a,th=4,tw=8):
        Hc,Wc=a.shape[0]//th*th,a.shape[1]//tw*tw
        return a[:Hc,:Wc].reshape(Hc//th,th,Wc//tw,tw).transpose(0,2,1,3).reshape(-1,th*tw)
    stats=dict(iters=0, box_now=0, blob_now=0, box_new=0, blob_new=0, lane_evals=0, lane_blob_now=0, lane_blob_new=0)
    def march(o, rd, act, maxit, far, shadow=False, ld=None):
        t=np.zeros(act.shape,f32); res=np.ones(act.shape,f32)
        for i in range(maxit):
            p=o+rd*t[...,None]
            plane=p[...,1]+1
            bs=boxskip(p,plane)            # today's decision (lane)
            # new: far-check first
            u=plane*f32(1.004)+f32(0.01)
            fskip=far>u
            need_box=act&~fskip
            need_blob_new=need_box&~bs
            need_blob_now=act&~bs
            b=blob(p)
            d=np.where(bs|fskip, plane, np.minimum(plane,b))
            # exactness sanity: skipping must not change the min
            assert (np.minimum(plane,b)[act&(bs|fskip)]==plane[act&(bs|fskip)]).all()
            ta=tiles(act).any(1)
            stats['iters']+=ta.sum(); stats['box_now']+=ta.sum(); stats['blob_now']+=tiles(need_blob_now).any(1).sum()
            stats['box_new']+=tiles(need_box).any(1).sum(); stats['blob_new']+=tiles(need_blob_new).any(1).sum()
            stats['lane_evals']+=act.sum(); stats['lane_blob_now']+=need_blob_now.sum(); stats['lane_blob_new']+=need_blob_new.sum()
            # update far: evaluated -> blob distance; else keep.  then decrement by the step
            far=np.where(need_blob_new, b, far)
            d=np.where(act,d,0)
            if shadow:
                with np.errstate(all='ignore'): q=50*d/t
                res=np.where(act,np.fmin(res,q),res)
            t=t+d
            far=far-np.abs(d)*f32(1.0001)
            if shadow: act=act&~((res<-1)|(t>ld)|(res<=0))
            else: act=act&~((d<1e-3)|(t>100))
            if not act.any(): break
        return t, far
    cam=np.array([-2,6,3],f32); dd=np.array([.3,-.7,-1],f32); fov=150
    dirn=dd/np.linalg.norm(dd); hh=np.arctan(np.radians(fov)/2); ww=W/H*hh
    right=np.cross(dirn,[0,1,0]); right/=np.linalg.norm(right); up=np.cross(right,dirn)
    xs=(np.arange(W)+.5)/W*2-1; ys=1-(np.arange(H)+.5)/H*2
    vx,vy=np.meshgrid(xs,ys)
    rd=right*(vx*ww)[...,None]+up*(vy*hh)[...,None]+dirn; rd/=np.linalg.norm(rd,axis=-1,keepdims=True); rd=rd.astype(f32)
    far0=np.full((H,W),-np.inf,f32)
    t,far=march(cam,rd,np.ones((H,W),bool),256,far0)
    print("primary:",{k:int(v) for k,v in stats.items()})
    hit=t<100
    p=cam+rd*t[...,None]
    h=(t/100)[...,None]
    ks=np.array([[1,-1,-1],[-1,-1,1],[-1,1,-1],[1,1,1]],f32)
    def sdf(p): return np.minimum(blob(p),p[...,1]+1)
    n=sum(kk*sdf(p+kk*h)[...,None] for kk in ks); n/=np.linalg.norm(n,axis=-1,keepdims=True)
    L=np.array([[-2,10,-1],[-7,2,-5]],f32)
    for li in range(2):
        l=L[li]-p; ld=np.linalg.norm(l,axis=-1); l/=ld[...,None]
        ndl=(n*l).sum(-1)
        act=hit&(ndl>0)
        march(p+l, l, act, 128, far-f32(1.001), shadow=True, ld=ld)
    print("all:",{k:int(v) for k,v in stats.items()})
    s=stats
    now=s['iters']*(13+4+6+16)+s['blob_now']*110
    new=s['iters']*(13+4+6+4)+s['box_new']*16+s['blob_new']*110
    print("warp-instr model: now",now,"new",new,"ratio",new/now, " blob evals warp-level now/new", s['blob_now']/s['iters'], s['blob_new']/s['iters'], "box tests new", s['box_new']/s['iters'])
