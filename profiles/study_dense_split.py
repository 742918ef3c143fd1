"""Design study: split each row block's column atoms into 'near' (guaranteed inside the cutoff of EVERY row atom) and the rest."""
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
rows_sel = rng.choice(nb-1, 60, replace=False)
for alpha in (1.0, 0.8, 0.6, 0.4):
    tot_hits=0; near_hits=0; near_slots=0; tiles_near=0; tiles_far=0; far_hits=0
    for rb in rows_sel:
        ra = np.arange(rb*32, rb*32+32)
        p = xs[ra]; ref = p[0]; q = p - L*np.round((p-ref)/L); ctr = 0.5*(q.max(0)+q.min(0))  # bbox centre as in the kernel
        R = np.linalg.norm(q-ctr,axis=1).max()
        nbrs = tree.query_ball_point(xs[ra], cutoff+pad)
        cols = np.unique(np.concatenate([np.array(n) for n in nbrs])); cols = cols[cols >= rb*32+32]
        hit = tree.query_ball_point(xs[ra], cutoff)
        d = xs[cols]-ctr; d -= L*np.round(d/L); r = np.linalg.norm(d,axis=1)
        near = r < cutoff - alpha*R
        hs = np.zeros(len(cols), int); cidx={c:k for k,c in enumerate(cols)}
        for i,h in enumerate(hit):
            for c in h:
                if c in cidx: hs[cidx[c]] += 1
        tot_hits += hs.sum(); near_hits += hs[near].sum(); near_slots += 32*near.sum(); far_hits += hs[~near].sum()
        tiles_near += int(np.ceil(near.sum()/32)); tiles_far += int(np.ceil((~near).sum()/32))
    print(f"alpha {alpha}: near columns hold {near_hits/tot_hits:.2%} of hits at fill {near_hits/max(near_slots,1):.2%}; tiles near {tiles_near} far {tiles_far} (far fill {far_hits/(tiles_far*1024):.2%}); block radius ~{R:.2f}")
