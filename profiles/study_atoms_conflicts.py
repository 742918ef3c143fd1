"""Design study: wavefronts per ATOMS instruction in phase B of the tile kernel for different accumulator layouts."""
import sys, numpy as np
sys.path.insert(0, '.')
from tests.common import water_box
from oracle import tm_oracle as O
from scipy.spatial import cKDTree
s = water_box(10000, seed=2022, jitter=0.01)
x, box = s['x'], s['box']; L = box[0,0]; N = len(x); cutoff, pad = 1.2, 0.1
perm = O.hilbert_perm(x, box); xs = x[perm] % L
tree = cKDTree(xs, boxsize=L)
rng = np.random.default_rng(0)
nb = (N+31)//32
def passes(banks):
    return np.bincount(banks, minlength=32).max()
res = {}
def add(k, v): res.setdefault(k, []).append(v)
for rb in rng.choice(nb-1, 40, replace=False):
    ra = np.arange(rb*32, rb*32+32)
    nbrs = tree.query_ball_point(xs[ra], cutoff+pad)
    cols = np.unique(np.concatenate([np.array(n) for n in nbrs])); cols = cols[cols >= rb*32+32]  # off-diagonal tiles only
    hit = tree.query_ball_point(xs[ra], cutoff)
    cidx = {c:k for k,c in enumerate(cols)}
    ntile = (len(cols)+31)//32
    H = np.zeros((32, ntile*32), bool)
    for i,h in enumerate(hit):
        for c in h:
            if c in cidx: H[i, cidx[c]] = True
    for t in range(ntile):
        Ht = H[:, t*32:(t+1)*32]
        # queue order: double round R (cols i+2R, i+2R+1): first all lanes' first column hits (b0), then second (b1)
        q = []
        for R in range(16):
            for sub in range(2):
                for i in range(32):
                    j = (i + 2*R + sub) % 32
                    if Ht[i, j]: q.append((i, j))
        q = np.array(q).reshape(-1, 2)
        for b0 in range(0, len(q) - 31, 32):
            b = q[b0:b0+32]; lane = np.arange(32)
            add('row 1copy', passes(b[:,0])); add('col 1copy', passes(b[:,1]))
            add('row 2copy lane&1 +16', passes((b[:,0] + 16*(lane&1)) % 32))
            add('col 2copy lane&1 +16', passes((b[:,1] + 16*(lane&1)) % 32))
            add('row 2copy lane>>4 +16', passes((b[:,0] + 16*(lane>>4)) % 32))
            add('col 2copy lane>>4 +16', passes((b[:,1] + 16*(lane>>4)) % 32))
            add('row 4copy lane&3 +8', passes((b[:,0] + 8*(lane&3)) % 32))
            add('col 4copy lane&3 +8', passes((b[:,1] + 8*(lane&3)) % 32))
            add('row 2copy +1', passes((b[:,0] + 1*(lane&1)) % 32))
            # ideal: distinct addresses spread perfectly = multiplicity bound
            add('row distinct-address floor', max(1, int(np.ceil(32/ max(1,len(set(b[:,0])))))) )
for k,v in res.items(): print(f"{k:32s} {np.mean(v):.2f}")
