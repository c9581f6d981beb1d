"""Development helper (CPU, numpy): a model of the fused assembly plan of the P1 tetrahedral workload that counts the
shared-memory wavefronts of phase 2 for candidate layouts of the local matrices, without a GPU.

It rebuilds, for a unit-cube Kuhn mesh, what csrc/pattern.cu builds on the device -- contributions sorted by (row, col,
cell, slot), rows in Morton order, blocks of 32 rows, ascending cell lists, entries ordered by descending segment length --
and then replays phase 2 warp by warp: a 64-bit shared-memory load of a warp is two half-warp passes, and inside a pass
distinct 8-byte words that fall into the same bank pair (word index mod 16) serialise.

    python tools/plan_model.py [n]        (default n = 32: 196,608 tets)

Result (n = 32, 60 sampled blocks): 5.13 wavefronts per load for the layout used by the kernel (loc[slot * lcap + cell]) --
ncu measures 5.2 on the real C4 run (2233196 wavefronts / 429493 instructions), so the model is faithful -- and no
plain re-layout helps: padding the slot stride 5.25, cell-major with stride 11: 5.21, cell-major stride 10: 6.98.  Sixteen
lanes reading effectively random cells hit sixteen bank pairs like balls into bins (expected maximum ~3 per pass).
A conflict-aware numbering does better: 'windowed_banks' (what csrc/pattern.cu:k_bank_colour builds, opt-in with
FDB_FUSED_BANKS=1) keeps every cell inside its 16-cell window of the list and picks its position greedily against its
co-readers: 3.94 wavefronts per load.  Not part of the product."""
import os
import sys

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
fdb=g.load_package()
n=int(sys.argv[1]) if len(sys.argv)>1 else 32
nodes,cells,bnd=fdb.meshes.unit_cube(n)
nc=cells.shape[0]; nn=nodes.shape[0]
# symmetric emission: pairs a<=b, row=max dof, col=min dof; contributions sorted by (row,col,cell,slot)
pairs=[(a,b) for a in range(4) for b in range(a,4)]
rows=np.empty((nc,10),np.int64); cols=np.empty((nc,10),np.int64)
for s,(a,b) in enumerate(pairs):
    da,db=cells[:,a].astype(np.int64),cells[:,b].astype(np.int64)
    rows[:,s]=np.maximum(da,db); cols[:,s]=np.minimum(da,db)
key=(rows*nn+cols).ravel()
cell_of=np.repeat(np.arange(nc),10); slot_of=np.tile(np.arange(10),nc)
order=np.argsort(key,kind='stable')
key_s=key[order]; cell_s=cell_of[order]; slot_s=slot_of[order]
heads=np.r_[True,key_s[1:]!=key_s[:-1]]
uid=np.cumsum(heads)-1; nu=uid[-1]+1
seg=np.r_[np.nonzero(heads)[0],key_s.size]
urow=(key_s[heads]//nn)
# Morton order of rows by centroid of an incident cell (use the row's own node coords as proxy)
def spread(x):
    x=x.astype(np.uint64)&0x1fffff
    x=(x|x<<32)&0x1f00000000ffff; x=(x|x<<16)&0x1f0000ff0000ff; x=(x|x<<8)&0x100f00f00f00f00f
    x=(x|x<<4)&0x10c30c30c30c30c3; x=(x|x<<2)&0x1249249249249249
    return x
# incident cell = first contribution of the row's last entry
last_u=np.r_[np.nonzero(urow[1:]!=urow[:-1])[0],nu-1]
row_ids=urow[last_u]
ecell=cell_s[seg[last_u]]
cent=nodes[cells[ecell]].mean(axis=1)
q=(cent*2097151).astype(np.int64)
mkey=spread(q[:,0])|(spread(q[:,1])<<np.uint64(1))|(spread(q[:,2])<<np.uint64(2))
rorder=row_ids[np.argsort(mkey,kind='stable')]
rank=np.empty(nn,np.int64); rank[rorder]=np.arange(nn)
rb=32
blk_of_u=rank[urow]//rb
nblocks=(nn+rb-1)//rb
print('n',n,'cells',nc,'unique',nu,'blocks',nblocks)

# ---- replay of phase 2 ---------------------------------------------------------------------------------------------
nu=blk_of_u.size
lens=np.diff(seg)
NT=224
def wavefronts64(addr_words):
    # addr_words: array of 8-byte word indices accessed by up to 32 lanes (-1 = inactive). 64-bit access: two half-warps,
    # each half: banks = word % 16; distinct words in the same bank serialize
    w=0
    for h in (addr_words[:16], addr_words[16:]):
        h=h[h>=0]
        if h.size==0: continue
        uniq=np.unique(h)
        banks=uniq%16
        w+=np.bincount(banks,minlength=16).max()
    return w
def simulate(layout, sample_blocks):
    tot_w=0; tot_i=0; rho_cells=0
    for b in sample_blocks:
        us=np.nonzero(blk_of_u==b)[0]
        if us.size==0: continue
        # block cell list (ascending) and local numbering
        contrib_idx=np.concatenate([np.arange(seg[u],seg[u+1]) for u in us])
        bcells=np.unique(cell_s[contrib_idx])
        lcap=(bcells.size+31)//32*32
        lc_of={c:i for i,c in enumerate(bcells)}
        # entries sorted by descending length, stable in u
        o=np.argsort(-lens[us],kind='stable'); us_o=us[o]
        # per entry list of word addresses
        def addr(c,s):
            lc=lc_of[c]
            if layout=='slot_major': return s*lcap+lc
            if layout=='cell_major11': return lc*11+s
            if layout=='slot_major_pad1': return s*(lcap+1)+lc
            if layout=='cell_major10': return lc*10+s
        if layout=='windowed_banks':
            # csrc/pattern.cu:k_bank_colour -- cells keep their 16-cell window, the position inside it is chosen greedily
            # against the cells read in the same (half-warp, step) group
            raw=[[(cell_s[t],slot_s[t]) for t in range(seg[u],seg[u+1])] for u in us_o]
            groups_of={}
            lmax=max(len(l) for l in raw)
            for k,l in enumerate(raw):
                for j,(c,_) in enumerate(l): groups_of.setdefault(c,[]).append((k>>4)*lmax+j)
            hist=np.zeros((((len(raw)+15)>>4)*lmax,16),int)
            newpos={}
            for w0 in range(0,bcells.size,16):
                win=bcells[w0:w0+16]; free=list(range(len(win)))
                for c in win:
                    gs=groups_of.get(c,[])
                    cost=[hist[gs,k].sum() for k in free]
                    k=free[int(np.argmin(cost))]; free.remove(k)
                    newpos[c]=w0+k
                    for gi in gs: hist[gi,k]+=1
            lists=[[s_*lcap+newpos[c] for c,s_ in l] for l in raw]
        else:
            lists=[[addr(cell_s[t],slot_s[t]) for t in range(seg[u],seg[u+1])] for u in us_o]
        ne=len(lists)
        for k0 in range(0,ne,32):
            grp=lists[k0:k0+32]
            L=max(len(l) for l in grp)
            for j in range(L):
                a=np.full(32,-1,np.int64)
                for li,l in enumerate(grp):
                    if j<len(l): a[li]=l[j]
                tot_w+=wavefronts64(a); tot_i+=1
    return tot_w/tot_i, tot_i
rng=np.random.default_rng(0)
sample=rng.choice(nblocks,60,replace=False)
for layout in ('slot_major','slot_major_pad1','cell_major11','cell_major10','windowed_banks'):
    print(layout, simulate(layout,sample))
