"""Cost model of decode_write_kernel's phase loop over the bench corpus' symbols-per-block distribution:
fixed phase lengths against an adaptive rule. CPU only; prints cost relative to 5 steps per phase."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from jpeg_rust_b200 import synth
rng=np.random.default_rng(0)
# symbols per block distribution from a real image (interleaved order irrelevant here)
data, coefs = synth.synth_jpeg(3, 1920, 1080, "420", want_coefs=True)
lens=[]
for c,comp in enumerate(coefs):
    for blk in comp:
        nz=np.nonzero(blk[1:])[0]+1
        n=1; last=0
        for pos in nz:
            run=pos-last-1; n+= run//16 + 1; last=pos
        if last<63: n+=1
        lens.append(n)
lens=np.array(lens); print("mean symbols/block", lens.mean())
STEP=55.0; BLOCKEND=36.0; PHASE_FIXED=20.0; FLUSH_PER4=12.0
def run(policy, P, thresh=None, nwarps=200, nblocks=160):
    tot_cost=0.0; tot_sym=0
    for w in range(nwarps):
        seqs=[rng.choice(lens, nblocks) for _ in range(32)]
        idx=[0]*32; rem=[int(s[0]) for s in seqs]; done=[False]*32
        cost=0.0
        while not all(done):
            blocked=[False]*32; k=0
            while True:
                act=[i for i in range(32) if not done[i] and not blocked[i]]
                if not act: break
                cost+=STEP   # one warp step (all active lanes in lockstep)
                for i in act:
                    rem[i]-=1; tot_sym+=1
                    if rem[i]==0: blocked[i]=True
                k+=1
                nb=sum(blocked)
                if policy=='fixed' and k>=P: break
                if policy=='adaptive' and (k>=P or nb>=thresh): break
            nb=sum(blocked)
            cost+=PHASE_FIXED + (BLOCKEND if nb else 0) + FLUSH_PER4*((nb+3)//4)
            for i in range(32):
                if blocked[i]:
                    idx[i]+=1
                    if idx[i]>=nblocks: done[i]=True
                    else: rem[i]=int(seqs[i][idx[i]])
        tot_cost+=cost
    return tot_cost/ (tot_sym/32.0)   # warp-instr per warp-symbol-slot
base=run('fixed',5)
print("fixed P=5", round(base,1))
for P in (3,4,6,8): print("fixed P",P, round(run('fixed',P)/base,3))
for P in (6,8,12):
    for th in (6,8,10,12,16): print("adaptive P",P,"thresh",th, round(run('adaptive',P,th)/base,3))
