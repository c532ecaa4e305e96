"""Seeded synthetic YouCookII-shaped inputs: conv5 maps, RPN-like
proposals, query lengths.  numpy only (RandomState: bit-stable across numpy versions)."""
import numpy as np

# histogram of entities per segment in the reference's train_entities.pkl for 0..13 entities
# (probe of data/YouCookII)
ENTITY_LEN_HIST = [1120, 2908, 3166, 1665, 813, 355, 168, 72, 39, 20, 6, 3, 1, 1]


def conv5_maps(rs, F, C, H, W):
    """relu(N(0,1)) like VGG16 conv5_3+ReLU (vgg16_rpn.py:38 keeps the ReLU, drops the pool)."""
    return np.maximum(rs.standard_normal((F, C, H, W)).astype(np.float32), 0.0)


def proposals(rs, F, n, img_h, img_w, cluster_frac=0.8):
    """Per frame n boxes + scores, score-descending (the NMS input contract).

    Base boxes: centre ~U(image), side 16*2^U(0, log2(min(H,W)/16)), aspect in {1/2,1,2}*U(.8,1.25).
    A fraction `cluster_frac` are jittered copies of one of ~n/96 randomly placed seed boxes, the
    way RPN anchors at neighbouring cells regress to the same object, so NMS has real work to do
    from the very first boxes on.
    Clipped to [0, W-1] x [0, H-1] (bbox_transform.py:125-133).
    """
    out = np.zeros((F, n, 4), np.float32)
    for f in range(F):
        cx = rs.uniform(0, img_w, n)
        cy = rs.uniform(0, img_h, n)
        side = 16.0 * 2.0 ** rs.uniform(0, np.log2(min(img_h, img_w) / 16.0), n)
        asp = rs.choice([0.5, 1.0, 2.0], n) * rs.uniform(0.8, 1.25, n)
        w = side * np.sqrt(asp)
        h = side / np.sqrt(asp)
        n_seed = max(1, n // 96)
        seeds = rs.permutation(n)[:n_seed]
        clustered = rs.uniform(0, 1, n) < cluster_frac
        clustered[seeds] = False
        src = seeds[rs.randint(0, n_seed, n)]
        jit = rs.standard_normal((n, 4))
        cx = np.where(clustered, cx[src] + 0.08 * w[src] * jit[:, 0], cx)
        cy = np.where(clustered, cy[src] + 0.08 * h[src] * jit[:, 1], cy)
        w = np.where(clustered, w[src] * np.exp(0.08 * jit[:, 2]), w)
        h = np.where(clustered, h[src] * np.exp(0.08 * jit[:, 3]), h)
        b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
        b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
        out[f] = b.astype(np.float32)
    scores = -np.sort(-rs.uniform(0, 1, (F, n)).astype(np.float32), axis=1)
    return out, scores


def entity_lengths(rs, Na, Ne):
    """Query counts per segment from the train histogram; resampled while the whole batch is
    empty (the reference skips such batches, model.py:685)."""
    p = np.asarray(ENTITY_LEN_HIST[: Ne + 1], np.float64)
    p /= p.sum()
    while True:
        lens = rs.choice(len(p), Na, p=p)
        if lens.sum() > 0:
            return [int(x) for x in lens]


def embeddings(rs, rows, D):
    """tanh-bounded rows like VisEbd / WordEbd outputs (model.py:616-642); IEEE-exact recipe."""
    return np.clip(rs.standard_normal((rows, D)) * 0.5, -1, 1).astype(np.float32)


CONFIGS = {
    # BASELINE.json configs[0]: one segment, CPU-runnable, eval phase
    "cfg1": dict(Na=1, Ns=5, Nb=20, Ne=13, D=512, H=38, W=50, C=512, n=2352, pre=6000,
                 img_h=608, img_w=800, train=False, Delta=5.0, vis_lam=1.0),
    # configs[1]: the training step the metric is quoted on
    "cfg2": dict(Na=8, Ns=5, Nb=20, Ne=13, D=512, H=38, W=50, C=512, n=2352, pre=6000,
                 img_h=608, img_w=800, train=True, Delta=10.0, vis_lam=4.13),
    # reference-real variant: 224x224 frames -> 14x14 conv5 maps (model.py:131-136)
    "cfg2_real": dict(Na=8, Ns=5, Nb=20, Ne=13, D=512, H=14, W=14, C=512, n=2352, pre=6000,
                      img_h=224, img_w=224, train=True, Delta=10.0, vis_lam=4.13),
    # configs[3]: dense-proposal stress
    "cfg4": dict(Na=1, Ns=32, Nb=100, Ne=13, D=512, H=38, W=50, C=512, n=300, pre=6000,
                 img_h=608, img_w=800, train=False, Delta=5.0, vis_lam=1.0),
    # configs[4]: inference sweep -- 10 000 cfg1-shaped segments (eval phase, batch_size_val = 1),
    # G segments per launch set, sharded contiguously over the ranks
    "cfg5": dict(Na=1, Ns=5, Nb=20, Ne=13, D=512, H=38, W=50, C=512, n=2352, pre=6000,
                 img_h=608, img_w=800, train=False, Delta=5.0, vis_lam=1.0, G=8, segments=10000,
                 queries=4, classes=67, pool=64),
}


def sweep_pool(c, seed=4242):
    """Host-side pieces of the cfg5 sweep that do not depend on the sharding: a pool of `pool`
    segments' proposals / scores (segment s uses entry s % pool), the class ids of every segment's
    queries (without replacement from `classes`, -1 in padded slots) and one ground-truth box per
    (segment, frame, query): a jittered copy of one of that frame's first 40 proposals."""
    rs = np.random.RandomState(seed)
    P, Ns, Ne, S, Q = c["pool"], c["Ns"], c["Ne"], c["segments"], c["queries"]
    props, scores = proposals(rs, P * Ns, c["n"], c["img_h"], c["img_w"])
    cls = np.full((S, Ne), -1, np.int32)
    for s in range(S):
        cls[s, :Q] = rs.choice(c["classes"], Q, replace=False)
    pick = rs.randint(0, 40, (S, Ns, Ne))
    base = props.reshape(P, Ns, c["n"], 4)[np.arange(S)[:, None, None] % P, np.arange(Ns)[None, :, None], pick]
    gt = base.astype(np.float64) + rs.uniform(-6, 6, (S, Ns, Ne, 4))
    gt[..., 2:] = np.maximum(gt[..., 2:], gt[..., :2] + 1)
    lens = np.full((S,), Q, np.int32)
    return dict(proposals=props.reshape(P, Ns, c["n"], 4), scores=scores.reshape(P, Ns, c["n"]),
                classes=cls, gt_boxes=gt, lens=lens)


def make_batch(cfg, seed):
    """One synthetic batch of the named config as a dict of numpy arrays."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    rs = np.random.RandomState(seed)
    F = c["Na"] * c["Ns"]
    props, scores = proposals(rs, F, c["n"], c["img_h"], c["img_w"])
    lens = entity_lengths(rs, c["Na"], c["Ne"]) if c["train"] else [min(4, c["Ne"])] * c["Na"]
    return dict(
        features=conv5_maps(rs, F, c["C"], c["H"], c["W"]),
        proposals=props,
        scores=scores,
        vis_feats=embeddings(rs, F * c["Nb"], c["D"]),
        word_feats=embeddings(rs, c["Na"] * c["Ne"], c["D"]),
        lens=lens,
    )
