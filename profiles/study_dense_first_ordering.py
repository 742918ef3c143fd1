"""Design study: ATOMS passes if the columns of a row block's tiles are ordered by distance to the row block (dense tiles first)."""
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
def passes(banks): return np.bincount(banks, minlength=32).max()
def study(order_fn, label):
    tot_hits=0; w_row=0; w_col=0; nb_batches=0; fills=[]; util=0
    for rb in rows_sel:
        ra = np.arange(rb*32, rb*32+32)
        nbrs = tree.query_ball_point(xs[ra], cutoff+pad)
        cols = np.unique(np.concatenate([np.array(n) for n in nbrs])); cols = cols[cols >= rb*32+32]
        hit = tree.query_ball_point(xs[ra], cutoff)
        cols = order_fn(cols, ra)
        cidx = {c:k for k,c in enumerate(cols)}
        ntile = (len(cols)+31)//32
        H = np.zeros((32, ntile*32), bool)
        for i,h in enumerate(hit):
            for c in h:
                if c in cidx: H[i, cidx[c]] = True
        for t in range(ntile):
            Ht = H[:, t*32:(t+1)*32]
            fills.append(Ht.mean())
            q = []
            for R in range(16):
                for sub in range(2):
                    for i in range(32):
                        j = (i + 2*R + sub) % 32
                        if Ht[i, j]: q.append((i, j))
            q = np.array(q).reshape(-1, 2)
            for b0 in range(0, len(q), 32):
                b = q[b0:b0+32]
                w_row += passes(b[:,0]); w_col += passes(b[:,1]); nb_batches += 1; util += len(b)
            tot_hits += len(q)
    print(f"{label:40s} tiles/row {ntile} batches {nb_batches} hits {tot_hits} lane util {util/(32*nb_batches):.2f} row passes/batch {w_row/nb_batches:.2f} col passes/batch {w_col/nb_batches:.2f} ATOMS wavefronts per hit {(6*(w_row+w_col))/tot_hits:.3f}")
    f=np.array(fills); print("    fill quantiles", np.quantile(f,[0.1,0.25,0.5,0.75,0.9]).round(2))
rows_sel = rng.choice(nb-1, 30, replace=False)
def as_is(cols, ra): return cols
def by_dist(cols, ra):
    ctr = xs[ra].mean(0)
    d = xs[cols]-ctr; d -= L*np.round(d/L)
    return cols[np.argsort(np.linalg.norm(d,axis=1))]
def by_dist_bins(nbins):
    def f(cols, ra):
        ctr = xs[ra].mean(0)
        d = xs[cols]-ctr; d -= L*np.round(d/L); r=np.linalg.norm(d,axis=1)
        b = np.minimum((r/ (1.9/nbins)).astype(int), nbins-1)
        return cols[np.argsort(b, kind='stable')]
    return f
study(as_is, "index order (now)")
study(by_dist, "sorted by distance to row block centre")
study(by_dist_bins(4), "4 distance bins")
study(by_dist_bins(8), "8 distance bins")
