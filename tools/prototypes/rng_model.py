"""Pure-Python model of the numpy Generator pieces the path consumes (SeedSequence, PCG64, buffered
uint32, Lemire bounded integers, random_interval, binomial n=1), checked against the installed numpy.
Development record for pogema_b200/csrc/pgm_rng.h.

    python tools/prototypes/rng_model.py
"""
import numpy as np
M64=(1<<64)-1; M128=(1<<128)-1; M32=(1<<32)-1
MULT=0x2360ed051fc65da44385df649fccf645
# SeedSequence
INIT_A=0x43b0d7e5; MULT_A=0x931e8875; INIT_B=0x8b51f9dd; MULT_B=0x58f38ded
MIX_MULT_L=0xca01f9dd; MIX_MULT_R=0x4973f715; XSHIFT=16
def seedseq_pool(entropy):
    # entropy int -> uint32 words little endian
    ws=[]
    e=entropy
    if e==0: ws=[0]
    while e>0:
        ws.append(e&M32); e>>=32
    pool=[0]*4
    hc=[INIT_A]
    def hashmix(v):
        v=(v^hc[0])&M32
        hc[0]=(hc[0]*MULT_A)&M32
        v=(v*hc[0])&M32
        v^=v>>XSHIFT
        return v
    def mix(x,y):
        r=((MIX_MULT_L*x)&M32)-((MIX_MULT_R*y)&M32)
        r&=M32
        r^=r>>XSHIFT
        return r
    for i in range(4):
        pool[i]=hashmix(ws[i] if i<len(ws) else 0)
    for i_src in range(4):
        for i_dst in range(4):
            if i_src!=i_dst:
                pool[i_dst]=mix(pool[i_dst],hashmix(pool[i_src]))
    for i_src in range(4,len(ws)):
        for i_dst in range(4):
            pool[i_dst]=mix(pool[i_dst],hashmix(ws[i_src]))
    return pool
def generate_state64(pool,n):
    hc=INIT_B
    out32=[]
    for i in range(n*2):
        d=pool[i%4]
        d^=hc
        hc=(hc*MULT_B)&M32
        d=(d*hc)&M32
        d^=d>>XSHIFT
        out32.append(d)
    return [out32[2*i]|(out32[2*i+1]<<32) for i in range(n)]
class PCG64:
    def __init__(self,seed):
        s=generate_state64(seedseq_pool(seed),4)
        initstate=(s[0]<<64)|s[1]; initseq=(s[2]<<64)|s[3]
        self.state=0; self.inc=((initseq<<1)|1)&M128
        self.step(); self.state=(self.state+initstate)&M128; self.step()
        self.has32=0; self.u32=0
    def step(self): self.state=(self.state*MULT+self.inc)&M128
    def next64(self):
        self.step()
        hi=self.state>>64; lo=self.state&M64
        x=hi^lo; rot=hi>>58
        return ((x>>rot)|(x<<((-rot)&63)))&M64
    def next32(self):
        if self.has32:
            self.has32=0; return self.u32
        n=self.next64(); self.has32=1; self.u32=n>>32; return n&M32
    def next_double(self): return (self.next64()>>11)*(1.0/9007199254740992.0)
    def lemire32(self,rng):
        if rng==0: return 0
        rng_excl=rng+1
        m=self.next32()*rng_excl; left=m&M32
        if left<rng_excl:
            thr=(M32-rng)%rng_excl
            while left<thr:
                m=self.next32()*rng_excl; left=m&M32
        return m>>32
    def interval(self,mx):
        if mx==0: return 0
        mask=mx
        for s in (1,2,4,8,16,32): mask|=mask>>s
        if mx<=M32:
            while True:
                v=self.next32()&mask
                if v<=mx: return v
        while True:
            v=self.next64()&mask
            if v<=mx: return v
if __name__=='__main__':
    import random
    for seed in [0,1,42,2**31-2,2**40+17,2**63-5,12345678901234567890123]:
        g=np.random.default_rng(seed); m=PCG64(seed)
        st=g.bit_generator.state['state']
        assert st['state']==m.state and st['inc']==m.inc,(seed)
        a=g.integers(0,2**64,size=5,dtype=np.uint64); b=[m.next64() for _ in range(5)]
        assert list(map(int,a))==b
    # choice
    for seed in range(200):
        g=np.random.default_rng(seed); m=PCG64(seed)
        for t in range(20):
            K=random.randint(1,3000)
            comp=[(i,i+1) for i in range(K)]
            c=tuple(*g.choice(comp,1)); j=m.lemire32(K-1)
            assert int(c[0])==j,(seed,t,K,c,j)
    # integers int32max
    for seed in range(50):
        g=np.random.default_rng(seed); m=PCG64(seed)
        a=g.integers(np.iinfo(np.int32).max,size=7); b=[m.lemire32(2**31-2) for _ in range(7)]
        assert list(map(int,a))==b
        a=g.integers(5,size=9); b=[m.lemire32(4) for _ in range(9)]
        assert list(map(int,a))==b
    # shuffle list of tuples
    for seed in range(50):
        n=random.randint(1,800)
        order=[(i//30,i%30) for i in range(n)]; o2=list(order)
        np.random.default_rng(seed).shuffle(order)
        m=PCG64(seed)
        for i in range(n-1,0,-1):
            j=m.interval(i); o2[i],o2[j]=o2[j],o2[i]
        assert order==o2,(seed,n)
    # binomial
    import math
    for seed in range(50):
        for p in [0.0,0.1,0.3,0.5,0.7,0.95,1.0]:
            a=np.random.default_rng(seed).binomial(1,p,(16,16))
            m=PCG64(seed)
            if p==0.0: b=np.zeros((16,16),int)
            elif p<=0.5:
                q=1-p; qn=math.exp(1*math.log(q))
                b=np.array([1 if m.next_double()>qn else 0 for _ in range(256)]).reshape(16,16)
            else:
                pp=1-p  # numpy uses q=1-p as p
                if pp==0.0: b=np.ones((16,16),int)
                else:
                    q=1-pp; qn=math.exp(math.log(q))
                    b=np.array([0 if m.next_double()>qn else 1 for _ in range(256)]).reshape(16,16)
            assert (a==b).all(),(seed,p)
    print("all rng model checks pass")
