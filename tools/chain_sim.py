"""chain_sim.py -- CPU model of the full-label hash region: sectors touched per hit / per miss for slot
layouts and region sizes, with the table's own hash (grimb_group.h hash_key) on bench.py's 1M-haplotype
table and a probe stream of recombinants (12 % present).  Design aid for the probe kernel (DESIGN.md
section 8); nothing here runs on the product path.

    python tools/chain_sim.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
M32=np.uint64(0xFFFFFFFF)
def hash_key(k):
    k=k.astype(np.uint64)
    h=(k & M32) ^ (((k>>np.uint64(32))*np.uint64(0x9E3779B1)) & M32)
    h^=h>>np.uint64(16); h=(h*np.uint64(0x85ebca6b))&M32; h^=h>>np.uint64(13); h=(h*np.uint64(0xc2b2ae35))&M32; h^=h>>np.uint64(16)
    return h
names, fa, ff = bench.make_table(1000000)
bits=[10,11,10,8,10]; shift=np.cumsum([0]+bits[:-1])
keys=np.zeros(len(fa),np.uint64)
for l in range(5): keys |= fa[:,l].astype(np.uint64) << np.uint64(int(shift[l]))
rng=np.random.RandomState(1)
# miss keys: recombinants of random pairs (like the probe stream): mix loci of two haplotypes
p=ff[:,0]/ff[:,0].sum(); idx=rng.choice(len(p),size=(200000,2),p=p)
flip=rng.rand(200000,5)<0.5
rec=np.where(flip,fa[idx[:,0]],fa[idx[:,1]])
rk=np.zeros(len(rec),np.uint64)
for l in range(5): rk |= rec[:,l].astype(np.uint64) << np.uint64(int(shift[l]))
present=np.isin(rk,keys)
def simulate(slots_per_sector, n_sectors):
    # insert keys with linear probing by sector; return avg sectors touched per hit and per miss
    mask=n_sectors-1
    home=(hash_key(keys)&np.uint64(mask)).astype(np.int64)
    fill=np.zeros(n_sectors,np.int32)
    place=np.zeros(len(keys),np.int64)
    order=np.arange(len(keys))
    cur=home.copy(); pending=order
    # iterative placement: process in rounds (approximate insertion order effects ignored for chain statistics)
    # exact sequential insert is slow in python; use vectorised rounds: each round, keys try their current sector; sectors accept up to capacity
    steps=np.zeros(len(keys),np.int32)
    while len(pending):
        c=cur[pending]
        # rank within same sector among pending
        o=np.argsort(c,kind='stable'); cs=c[o]
        first=np.r_[0,np.flatnonzero(cs[1:]!=cs[:-1])+1]
        rank=np.arange(len(cs))-np.repeat(first,np.diff(np.r_[first,len(cs)]))
        free=slots_per_sector-fill[cs]
        ok=rank<free
        acc=pending[o][ok]
        np.add.at(fill,cs[ok],1)
        rej=pending[o][~ok]
        steps[rej]+=1
        cur[rej]=(cur[rej]+1)&mask
        pending=rej
    hit_sectors=1+steps.mean()
    # misses: walk from home until a sector with a free slot
    mh=(hash_key(rk[~present])&np.uint64(mask)).astype(np.int64)
    s=np.ones(len(mh)); c=mh.copy(); alive=fill[c]>=slots_per_sector
    while alive.any():
        c[alive]=(c[alive]+1)&mask; s[alive]+=1
        alive=alive&(fill[c]>=slots_per_sector)
    return hit_sectors, s.mean(), (fill>=slots_per_sector).mean()
print("recombinant probes present in table: %.3f"%present.mean())
for name,sps,nsec in [("16B slots x2, 128 MB (now)",2,1<<22),("16B slots x2, 64 MB",2,1<<21),("16B slots x2, 32 MB",2,1<<20),
                      ("8B slots x4, 64 MB",4,1<<21),("8B slots x4, 32 MB",4,1<<20),("8B slots x4, 16 MB",4,1<<19)]:
    h,m,full=simulate(sps,nsec)
    print("%-28s sectors/hit %.3f  sectors/miss %.3f  full sectors %.3f"%(name,h,m,full))
