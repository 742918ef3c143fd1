"""Offline analysis: per-lane hit counts of the 32x32 tiles of a Hilbert-sorted water box (design study for the
register-queue tile kernel)."""
import sys, numpy as np
sys.path.insert(0, '.')
from tests.common import water_box
from oracle import tm_oracle as O
from scipy.spatial import cKDTree

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
s = water_box(nw, seed=2022, jitter=0.01)
x, box = s['x'], s['box']; L = box[0,0]
N = len(x); cutoff, pad = 1.2, 0.1
perm = O.hilbert_perm(x, box)
xs = x[perm] % L
nb = (N + 31)//32
# block bounds (no PBC subtlety: use per-block min/max of wrapped coords relative to first atom)
tree = cKDTree(xs, boxsize=L)
# tile list as the reference: row block rb; column atoms j > ... within cutoff+pad of ANY row atom's bounding box -> approx: within cutoff+pad of any row atom
rng = np.random.default_rng(0)
rows = rng.choice(nb-1, 60, replace=False)
util = {1:[],2:[],3:[],4:[],8:[],16:[],'run':[]}
fills=[]
for rb in rows:
    ra = np.arange(rb*32, min(N, rb*32+32))
    # neighbours within cutoff+pad of any row atom (trimmed list like k_compact_trim_atoms), upper triangular: j >= rb*32
    nbrs = tree.query_ball_point(xs[ra], cutoff+pad)
    cols = np.unique(np.concatenate([np.array(n) for n in nbrs]))
    cols = cols[cols >= rb*32]
    # actual hits within cutoff
    hit = tree.query_ball_point(xs[ra], cutoff)
    H = np.zeros((len(ra), len(cols)), bool)
    cidx = {c:k for k,c in enumerate(cols)}
    for i,h in enumerate(hit):
        for c in h:
            if c in cidx and c > ra[i]: H[i, cidx[c]] = True
    ntile = (len(cols)+31)//32
    Hp = np.zeros((32, ntile*32), bool); Hp[:len(ra), :len(cols)] = H
    per = Hp.reshape(32, ntile, 32).sum(-1)  # [lane, tile]
    fills.append(per.sum()/ (ntile*1024))
    for g in (1,2,3,4,8,16):
        for t0 in range(0, ntile - g + 1, g):
            c = per[:, t0:t0+g].sum(1)
            if c.max() > 0: util[g].append((c.mean(), c.max()))
    c = per.sum(1); util['run'].append((c.mean(), c.max()))
print('N', N, 'tiles/row', ntile, 'fill', np.mean(fills))
for g,v in util.items():
    v = np.array(v); print(g, 'util = sum(mean)/sum(max) =', v[:,0].sum()/v[:,1].sum(), ' mean', v[:,0].mean(), 'max', v[:,1].mean())
