"""CPU model of shared-memory wavefronts per node load and of 128-byte lines per leaf-record gather on the C3 forests, for the
breadth-first and the depth-first node order (hm_forest.cu, hm_forest_node_order).  Warps = 32 consecutive candidates of a
2^16-candidate sample sorted by the leaves of two key trees (the real pool is 256 x denser, so absolute numbers are
pessimistic).  Result: bfs 1.78 wavefronts per LDS.64 / 12.0 lines per gather, dfs 1.82 / 10.3 -- and the measurement on the
GPU agreed that the order does not matter (profiles/r02_exp_round_overhead.log, block 7).    python tools/node_layout_sim.py"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from bench_support import workloads
wl=workloads.RFWorkload(workloads.scenario_c3(), n_train=2000, seed=5)
N=1<<16
X=np.ascontiguousarray(wl.candidates_encoded(N, seed=6).T)
ests={o:wl.models[o].estimators_ for o in wl.objectives}
def walk_leaf(tree,X):
    left,right,feat,thr=tree.children_left,tree.children_right,tree.feature,tree.threshold
    cur=np.zeros(len(X),dtype=np.int64)
    while True:
        act=left[cur]!=-1
        if not act.any(): break
        f=np.where(act,feat[cur],0)
        gl=X[np.arange(len(X)),f]<=thr[cur]
        cur=np.where(act,np.where(gl,left[cur],right[cur]),cur)
    return cur
def dfs_rank(tree):
    leaves=np.nonzero(tree.children_left==-1)[0]; r=np.full(tree.node_count,-1); r[leaves]=np.arange(len(leaves)); return r,len(leaves)
o0,o1=wl.objectives
kA,kB=ests[o0][0].tree_,ests[o1][0].tree_
rA,nA=dfs_rank(kA); rB,nB=dfs_rank(kB)
key=rA[walk_leaf(kA,X)]*nB+rB[walk_leaf(kB,X)]
# the real pool is 256x denser: emulate coherence by sorting this sample (a pessimistic stand-in)
order=np.argsort(key,kind='stable')
Xs=X[order]
def positions(tree,dfs):
    left,right=tree.children_left,tree.children_right
    n=tree.node_count; pos=np.full(n,-1); pos[0]=0; cnt=1
    work=[0]; head=0
    while (work if dfs else head<len(work)):
        if dfs: nd=work.pop()
        else: nd=work[head]; head+=1
        l,r=left[nd],right[nd]
        if l!=-1:
            pos[l]=cnt; pos[r]=cnt+1; cnt+=2
            if dfs: work+= [r,l]
            else: work+=[l,r]
    # leaf ordinal in layout order
    leaves=np.nonzero(left==-1)[0]
    lo=np.full(n,-1); srt=leaves[np.argsort(pos[leaves])]; lo[srt]=np.arange(len(srt))
    return pos,lo
def sim(tree,dfs,W=512):
    left,right,feat,thr=tree.children_left,tree.children_right,tree.feature,tree.threshold
    pos,lo=positions(tree,dfs)
    wf=0; loads=0; lines=0; gathers=0
    for w in range(W):
        xs=Xs[w*32:(w+1)*32]
        cur=np.zeros(32,dtype=np.int64)
        while True:
            act=left[cur]!=-1
            if not act.any(): break
            f=np.where(act,feat[cur],0)
            gl=xs[np.arange(32),f]<=thr[cur]
            cur=np.where(act,np.where(gl,left[cur],right[cur]),cur)   # arrived lanes re-load their own leaf
            addr=np.unique(pos[cur])
            bp=addr%16
            wf+=np.bincount(bp,minlength=16).max(); loads+=1
        lines+=len(np.unique(lo[cur]//4)); gathers+=1
    return wf/loads, lines/gathers
trees=[ests[o0][k].tree_ for k in (3,50,120,200)]+[ests[o1][k].tree_ for k in (9,77,180)]
for dfs in (False,True):
    r=[sim(t,dfs) for t in trees]
    print('dfs' if dfs else 'bfs','wavefronts per LDS.64 %.3f   128-byte lines per leaf gather %.2f'%(np.mean([a for a,b in r]),np.mean([b for a,b in r])))
