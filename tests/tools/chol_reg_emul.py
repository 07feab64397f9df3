"""numpy emulation of the data flow of csrc/cholesky_reg.cu (which lane holds which column, the layout of the
published pivot row, pivot rule, rank cut) for both lane layouts; no GPU needed.  python tests/tools/chol_reg_emul.py"""
import numpy as np
def emul(K, LPR, rel_tol=1e-5, NG=13):
    n=K.shape[0]; W=16//LPR
    rows=((n+ (32//LPR)-1)//(32//LPR))*(32//LPR)
    l=np.zeros((rows,NG,LPR,W),np.float32)
    di=np.full(rows,-1,np.float32); di[:n]=np.diag(K)
    LT=np.full((n,n),np.nan,np.float32)
    floor=np.float32(rel_tol)*max(di.max(),0)
    best_arr=di.copy(); rank=0; done=False
    for g in range(NG):
        for jq in range(LPR):
            if done: break
            for js in range(W):
                if done: break
                j=16*g+W*jq+js
                if j>=n: done=True; break
                cand=np.maximum(best_arr,0); p=int(np.argmax(cand)); best=np.float32(cand[p])
                if not (best>floor) or not (best>0): done=True; break
                prow=np.zeros(16*NG,np.float32)
                for gg in range(g+1):
                    for q in range(LPR):
                        prow[16*gg+W*q:16*gg+W*q+W]=l[p,gg,q]
                newb=np.full(rows,-1,np.float32)
                for i in range(rows):
                    acc=np.float32(0)
                    for q in range(LPR):
                        for gg in range(g+1):
                            acc+=np.dot(l[i,gg,q],prow[16*gg+W*q:16*gg+W*q+W])
                    c=np.float32(0); nd=np.float32(-1)
                    if i<n and di[i]>=0:
                        rs=np.float32(1)/np.sqrt(best)
                        c=best*rs if i==p else (np.float32(K[p,i])-acc)*rs
                        nd=np.float32(-1) if i==p else max(di[i]-c*c,np.float32(0)); di[i]=nd
                    l[i,g,jq,js]=c
                    if i<n: LT[j,i]=c; newb[i]=nd
                best_arr=newb; rank=j+1
    LT[rank:,:]=0
    return LT,rank
rng=np.random.default_rng(1)
for LPR in (2,4):
    for n,inner in [(196,384),(196,48),(205,400) if LPR==2 else (200,400)]:
        A=rng.standard_normal((n,inner))*np.logspace(0,-2,inner)
        K=(A@A.T).astype(np.float32)
        LT,rank=emul(K,LPR)
        err=np.abs(LT.T.astype(np.float64)@LT.astype(np.float64)-K).max()/np.abs(K).max()
        print(LPR,n,inner,'rank',rank,'err',err,'nan',np.isnan(LT).any())
