"""Prototype of the parallel (closed-form) move resolution used by the CUDA kernel (phase 2 of
pgm_step_kernel), fuzzed against the sequential oracle.  Development record: this is how the closed
forms in DESIGN.md section 3 were validated before they were written in CUDA.

    python tools/prototypes/closed_form.py 4000
"""
import sys, random
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np
from oracle.pogema_oracle import GridConfig, Pogema, MOVES

FAIL, OK, PEND = 0, 1, 2
def resolve(coll, obst, pos, active, act):
    A=len(pos)
    occ={}
    for i in range(A):
        if active[i]: occ[pos[i]]=i
    tgt=[(pos[i][0]+MOVES[act[i]][0], pos[i][1]+MOVES[act[i]][1]) for i in range(A)]
    def nbrs(c): return [(c[0]+dx,c[1]+dy) for dx,dy in MOVES[1:]]
    status=[FAIL]*A; parent=[-1]*A
    if coll=='block_both':
        for i in range(A):
            if not active[i] or act[i]==0: continue
            T=tgt[i]
            if obst[T]: continue
            if T in occ: continue
            ok=True
            for n in nbrs(T):
                if n==pos[i]: continue
                k=occ.get(n)
                if k is not None and act[k]!=0 and tgt[k]==T: ok=False
            status[i]=OK if ok else FAIL
    elif coll=='priority':
        for i in range(A):
            if not active[i] or act[i]==0: continue
            T=tgt[i]
            if obst[T]: continue
            j=occ.get(T)
            if j is not None and (j>i or act[j]==0): continue
            lo = j if j is not None else -1
            cand=i
            for n in nbrs(T):
                if n==pos[i]: continue
                k=occ.get(n)
                if k is not None and act[k]!=0 and tgt[k]==T and k>lo and k<cand: cand=k
            if cand!=i: continue
            if j is None: status[i]=OK
            else: status[i]=PEND; parent[i]=j
    else:
        eff=list(act)
        for i in range(A):
            if not active[i]: eff[i]=0; continue
            if act[i]==0: continue
            T=tgt[i]
            if obst[T]: eff[i]=0; continue
            j=occ.get(T)
            if j is not None and act[j]!=0 and tgt[j]==pos[i]: eff[i]=0
        for i in range(A):
            if not active[i] or eff[i]==0: continue
            T=tgt[i]
            j=occ.get(T)
            if j is not None and eff[j]==0: continue
            bad=False
            for n in nbrs(T):
                if n==pos[i]: continue
                k=occ.get(n)
                if k is not None and eff[k]!=0 and tgt[k]==T and k<i: bad=True
            if bad: continue
            if j is None: status[i]=OK
            else: status[i]=PEND; parent[i]=j
    # pointer jumping (synchronous rounds, reading old values)
    R=max(1,(A-1).bit_length())+1
    for _ in range(R):
        ns=list(status); npar=list(parent)
        for i in range(A):
            if status[i]==PEND:
                p=parent[i]
                if status[p]!=PEND: ns[i]=status[p]
                else: npar[i]=parent[p]
        status,parent=ns,npar
    out=[]
    for i in range(A):
        s=status[i]
        if s==PEND: s = OK  # only cycles remain (soft)
        out.append(tgt[i] if (s==OK and active[i] and act[i]!=0) else pos[i])
    return out

def fuzz(trials=3000, seed=0):
    rnd=random.Random(seed)
    for t in range(trials):
        H=rnd.randint(2,6); W=H
        dens=rnd.choice([0.0,0.1,0.3])
        cells=[(x,y) for x in range(H) for y in range(W)]
        grid=[[1 if rnd.random()<dens else 0 for _ in range(W)] for _ in range(H)]
        free=[c for c in cells if grid[c[0]][c[1]]==0]
        if len(free)<2: continue
        A=rnd.randint(1,min(len(free),12))
        starts=rnd.sample(free,A); goals=[rnd.choice(free) for _ in range(A)]
        for coll in ['priority','block_both','soft']:
            gc=GridConfig(map=grid,agents_xy=[list(s) for s in starts],targets_xy=[list(g) for g in goals],obs_radius=2,collision_system=coll,on_target='finish',seed=1)
            env=Pogema(gc); env.reset()
            for step in range(6):
                act=[rnd.randint(0,4) for _ in range(A)]
                g=env.grid
                pos=list(g.positions_xy); active=[g.is_active[i] for i in range(A)]
                obst={(x,y):bool(g.obstacles[x,y]) for x in range(g.obstacles.shape[0]) for y in range(g.obstacles.shape[1])}
                exp=resolve(coll,obst,pos,active,act)
                env.step(act)
                got=list(env.grid.positions_xy)
                assert got==exp,(t,coll,step,pos,act,active,got,exp)
                # occupancy consistency
                P=env.grid.positions
                occ=np.zeros_like(P)
                for i in range(A):
                    if env.grid.is_active[i]: occ[env.grid.positions_xy[i]]+=1
                assert (occ==P).all(),(t,coll,'occupancy')
    print('fuzz ok')
fuzz(int(sys.argv[1]) if len(sys.argv)>1 else 3000)
