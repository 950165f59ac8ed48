"""CPU simulation of the coherence pre-pass on the C3 forests: distinct nodes per warp visit, active lanes and distinct
leaves per warp for different sort keys (2^18-candidate sample; the real pool is 64x denser).  Result (round 1): a third
or fourth key tree changes nothing (2.50 -> 2.44 distinct nodes per warp visit): the residual divergence is inside a
bucket, not between buckets.

    python tools/coherence_sim.py
"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from bench_support import workloads
t=time.time()
wl=workloads.RFWorkload(workloads.scenario_c3(), n_train=2000, seed=5)
print('fit',time.time()-t)
N=1<<18
X=np.ascontiguousarray(wl.candidates_encoded(N, seed=6).T)   # [N][D] float32
ests={o:wl.models[o].estimators_ for o in wl.objectives}
def walk(tree, X):
    """returns list of node ids per level [L][N] (-1 after leaf) and leaf ids"""
    left,right,feat,thr=tree.children_left,tree.children_right,tree.feature,tree.threshold
    cur=np.zeros(len(X),dtype=np.int64); levels=[]
    active=np.ones(len(X),bool)
    while True:
        isleaf = left[cur]==-1
        active = ~isleaf
        if not active.any(): break
        levels.append(np.where(active,cur,-1))
        f=feat[cur]; go_left = X[np.arange(len(X)), np.where(active,f,0)] <= thr[cur]
        nxt=np.where(go_left,left[cur],right[cur])
        cur=np.where(active,nxt,cur)
    return levels,cur
def dfs_rank(tree):
    # leaf rank in DFS (node id order = preorder in sklearn)
    leaves=np.nonzero(tree.children_left==-1)[0]
    r=np.full(tree.node_count,-1); r[leaves]=np.arange(len(leaves)); return r,len(leaves)
o0,o1=wl.objectives
keyt=[ests[o0][0].tree_, ests[o1][0].tree_, ests[o0][1].tree_, ests[o1][1].tree_]
ranks=[]
for tr in keyt:
    _,leaf=walk(tr,X); r,n=dfs_rank(tr); ranks.append((r[leaf],n))
test_trees=[ests[o0][k].tree_ for k in (5,17,40,99,150,200)]+[ests[o1][k].tree_ for k in (7,33,120,210)]
paths=[walk(tr,X) for tr in test_trees]
def metric(order):
    tot_w=0; tot_dist=0; tot_lanes=0; leaf_dist=0; nw=0
    for levels,leaf in paths:
        lo=leaf[order].reshape(-1,32)
        leaf_dist+=sum(len(np.unique(r)) for r in lo[:2000]); nw+=2000
        for lv in levels:
            l=lv[order].reshape(-1,32)[:2000]
            act=(l>=0)
            anyact=act.any(1)
            for row,a in zip(l[anyact],act[anyact]):
                tot_w+=1; tot_dist+=len(np.unique(row[a])); tot_lanes+=a.sum()
    return tot_dist/tot_w, tot_lanes/tot_w, leaf_dist/nw
def order_of(keys):
    k=np.zeros(N,dtype=np.int64)
    for r,n,shift in keys: k=k*((n-1>>shift)+1)+(r>>shift)
    return np.argsort(k,kind='stable'), len(np.unique(k))
cfgs={'arrival':None,
 'A,B':[(ranks[0][0],ranks[0][1],0),(ranks[1][0],ranks[1][1],0)],
 'A,B,C>>4':[(ranks[0][0],ranks[0][1],0),(ranks[1][0],ranks[1][1],0),(ranks[2][0],ranks[2][1],4)],
 'A,B,C':[(ranks[0][0],ranks[0][1],0),(ranks[1][0],ranks[1][1],0),(ranks[2][0],ranks[2][1],0)],
 'A>>3,B>>3,C>>3,D>>3':[(ranks[i][0],ranks[i][1],3) for i in range(4)],
 'A>>2,B>>2,C>>2,D>>2':[(ranks[i][0],ranks[i][1],2) for i in range(4)],
}
for name,keys in cfgs.items():
    if keys is None: order=np.arange(N); nb=0
    else: order,nb=order_of(keys)
    d,l,ld=metric(order)
    print('%-24s bins=%8d distinct nodes/warp-visit %.2f  active lanes %.1f  distinct leaves/warp %.2f'%(name,nb,d,l,ld))
