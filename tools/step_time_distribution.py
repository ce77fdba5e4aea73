import sys; sys.path.insert(0,"sam-decoding_b200"); sys.path.insert(0,".")  # run from the repo root
import numpy as np, torch, bench
from samd_b200 import _cabi as K, engine as E
class A: requests=1024; prompt=8192; steps=32; warmup=8
streams, counts, tokens, start = bench.make_workload(1024, 8192, 40, 2000)
dev=torch.device("cuda")
dyn=E.DynSamBatch(1024, 8192+8*40+16, dev)
eng=E.DraftEngine(dyn, None, K.FLAVOUR_SAMD, n_predicts=16, len_bias=5, len_threshold=5)
eng.step(torch.as_tensor(streams[:, :8192]).to(dev), None, None)
dt,dc,ds=(torch.as_tensor(x).to(dev) for x in (tokens,counts,start))
cyc=torch.zeros(16,1024,dtype=torch.int64,device=dev)
K.lib().samd_step_set_debug_cycles(cyc.data_ptr())
allc=[]; feats=[]; phases=[]
import ctypes as C
def meta():
    m=np.zeros((1024,16),dtype=np.int32)
    K.check(K.lib().samd_dyn_meta(dyn.handle, m.ctypes.data_as(C.POINTER(C.c_int32))))
    return m.astype(np.int64)
m0=meta()
for s in range(40):
    eng.step(dt[s],dc[s],ds[s]); torch.cuda.synchronize()
    m1=meta()
    if s>=8:
        allc.append((cyc[0].cpu().numpy().copy(), counts[s].copy())); phases.append(cyc.cpu().numpy().copy())
        d=m1-m0
        feats.append(np.stack([np.ones(1024), counts[s], d[:,7], d[:,5], d[:,8], d[:,9]],1))   # 1, tokens, clones, edges, hops, probes
    m0=m1
K.lib().samd_step_set_debug_cycles(None)
c=np.stack([a for a,_ in allc]).astype(float)/1.9e3   # us at ~1.9GHz
k=np.stack([b for _,b in allc])
print("per-step max us: mean %.1f ; per-request mean %.1f p50 %.1f p90 %.1f p99 %.1f p99.9 %.1f" % (c.max(1).mean(), c.mean(), np.percentile(c,50), np.percentile(c,90), np.percentile(c,99), np.percentile(c,99.9)))
for kk in range(1,9):
    m=c[k==kk]; print("k=%d n=%d mean %.1f p99 %.1f max %.1f" % (kk, m.size, m.mean(), np.percentile(m,99), m.max()))
print("us per token (k>=1): %.2f" % (c.sum()/k.sum()))

# what a slow request is made of: least-squares fit of per-request time on the step's counters
X=np.concatenate(feats).astype(float); y=c.reshape(-1)
coef,*_=np.linalg.lstsq(X,y,rcond=None)
print("fit us = %.2f + %.2f*tokens + %.2f*clones + %.2f*edges + %.2f*link_hops + %.2f*probes ; residual sd %.2f" % (*coef, (y-X@coef).std()))
slow=y>np.percentile(y,99.5)
print("slowest 0.5%%: mean tokens %.1f clones %.1f edges %.1f hops %.1f probes %.1f (all: %.1f %.1f %.1f %.1f %.1f)" % (*X[slow,1:].mean(0), *X[:,1:].mean(0)))
order=np.argsort(-y)[:16]
print("slowest request-steps: time us | fitted | tokens clones edges hops probes")
for i in order:
    print("  %5.1f | %5.1f | %d %d %d %d %d" % (y[i], X[i]@coef, *X[i,1:].astype(int)))
ph=np.stack(phases).astype(float)/1.9e3            # [steps, 4, requests] us
tot,tr,ap,lk=ph[:,0].reshape(-1),ph[:,1].reshape(-1),ph[:,2].reshape(-1),ph[:,3].reshape(-1)
print("phase means us: total %.1f = cursor transfers %.1f + appends %.1f + lookup/draft %.1f + rest %.1f" % (tot.mean(), tr.mean(), ap.mean(), lk.mean(), (tot-tr-ap-lk).mean()))
sl=tot>np.percentile(tot,99.5)
print("slowest 0.5%%:    total %.1f = cursor transfers %.1f + appends %.1f + lookup/draft %.1f + rest %.1f" % (tot[sl].mean(), tr[sl].mean(), ap[sl].mean(), lk[sl].mean(), (tot-tr-ap-lk)[sl].mean()))
lkc,adc,miss=ph[:,4].reshape(-1),ph[:,5].reshape(-1),(ph[:,6].reshape(-1)*1.9e3)
print("inside the append loop (mean | slowest): look-ups %.2f | %.2f us, edge inserts %.2f | %.2f us" % (lkc.mean(), lkc[sl].mean(), adc.mean(), adc[sl].mean()))
qr,oc,rd=ph[:,7].reshape(-1),ph[:,8].reshape(-1),ph[:,9].reshape(-1)
print("at the chain's end (mean | slowest): target record %.2f | %.2f us, clone overflow copy %.2f | %.2f us, clone redirect walk %.2f | %.2f us" % (qr.mean(), qr[sl].mean(), oc.mean(), oc[sl].mean(), rd.mean(), rd[sl].mean()))
