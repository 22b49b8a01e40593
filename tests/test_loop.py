"""Loop-closure query slice (BASELINE configs[4]): the CPU oracle against hand-computed vectors
of DBoW2's formulas, and (gpu) svin_loop_* against the oracle - words, bag-of-words vectors, L1 scores, the top-4 and the
BRIEF candidate search bit-exact at small sizes, and at the 5k-keyframe size of configs[4]."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import loop_oracle as lo  # noqa: E402
from scene_loop import make_keyframes  # noqa: E402


def has_gpu():
    import torch
    return torch.cuda.is_available()


def tiny_vocabulary():
    """Root with two inner nodes of two leaves each; descriptors chosen so the descent is hand-checkable."""
    z = np.zeros(32, np.uint8)
    d = np.stack([z] * 7)
    d[1, :] = 0x00   # inner A: all zeros
    d[2, :] = 0xFF   # inner B: all ones
    d[3, :] = 0x00   # A0
    d[4, 0] = 0x0F   # A1: 4 bits set
    d[5, :] = 0xFF   # B0
    d[6, :] = 0xFF
    d[6, 0] = 0xF0   # B1: 4 bits cleared
    return lo.Vocabulary(first_child=[1, 3, 5, 0, 0, 0, 0], num_children=[2, 2, 2, 0, 0, 0, 0], descriptor=d,
                         weight=[0, 0, 0, 1.0, 2.0, 0.5, 0.0], word_id=[-1, -1, -1, 0, 1, 2, 3])


def feat(byte0, rest):
    f = np.full(32, rest, np.uint8)
    f[0] = byte0
    return f


# ---------------------------------------------------------------- oracle KATs (CPU)
def test_hamming_known_answers():
    a = np.zeros(32, np.uint8)
    b = np.full(32, 0xFF, np.uint8)
    assert lo.hamming(a, b) == 256 and lo.hamming(a, a) == 0
    c = a.copy()
    c[5] = 0b10110000
    assert lo.hamming(a, c) == 3


def test_descent_first_minimum_on_ties():
    v = tiny_vocabulary()
    assert v.transform_feature(feat(0x00, 0x00)) == (0, 1.0)
    assert v.transform_feature(feat(0x0F, 0x00)) == (1, 2.0)
    # 0x03: distance 2 to A0 and 2 to A1 -> the first child wins (TemplatedVocabulary.h:1137 `d < best_d`)
    assert v.transform_feature(feat(0x03, 0x00)) == (0, 1.0)
    # 128 bits set: tie at the root -> inner A
    half = np.zeros(32, np.uint8)
    half[:16] = 0xFF
    assert v.transform_feature(half)[0] in (0, 1)
    assert v.transform_feature(feat(0xF0, 0xFF)) == (3, 0.0)


def test_bow_tfidf_l1_normalised_and_zero_weight_words_dropped():
    v = tiny_vocabulary()
    fs = [feat(0x00, 0x00), feat(0x0F, 0x00), feat(0x00, 0x00), feat(0xFF, 0xFF), feat(0xF0, 0xFF)]
    ids, vals = v.transform(fs)
    # words 0 (twice, 1.0 each), 1 (2.0), 2 (0.5); word 3 has weight 0 and is not added (TemplatedVocabulary.h:1010)
    assert ids.tolist() == [0, 1, 2]
    np.testing.assert_array_equal(vals, np.array([2.0, 2.0, 0.5]) / 4.5)
    ids2, vals2 = v.transform(fs, fast=True)
    assert ids2.tolist() == ids.tolist() and (vals2 == vals).all()


def test_l1_score_known_answers():
    ids, a = np.array([0, 1, 2]), np.array([0.5, 0.25, 0.25])
    assert lo.l1_score(ids, a, ids, a) == 1.0                       # identical normalised vectors
    assert lo.l1_score(ids, a, np.array([3, 4]), np.array([0.5, 0.5])) == 0.0
    # one common word: -( |0.5-0.25| - 0.5 - 0.25 ) / 2 = 0.25
    assert lo.l1_score(ids, a, np.array([0, 7]), np.array([0.25, 0.75])) == 0.25


def test_database_query_order_max_id_and_cut():
    v = lo.Vocabulary.random(4, 3, seed=3)
    frames, _ = make_keyframes(24, 6, per_image=60, seed=5)
    db = lo.Database(v)
    for f in frames:
        db.add(f)
    r = db.query(frames[20], 4, -1)
    assert r[0][0] == 20 and r[0][1] == pytest.approx(1.0, abs=1e-12)
    assert [s for _, s in r] == sorted((s for _, s in r), reverse=True)
    r2 = db.query(frames[20], 4, 20)
    assert all(e < 20 for e, _ in r2) and r2[0][0] in (2, 8, 14)      # the earlier visits of place 2
    assert len(db.query(frames[20], 2, -1)) == 2
    fast = lo.Database(v, fast=True)
    for f in frames:
        fast.add(f)
    assert fast.query(frames[20], 4, 20) == r2


def test_detect_loop_decision():
    assert lo.detect_loop([(3, 0.5), (9, 0.2)], min_score=0.4, frame_index=100) == 3
    assert lo.detect_loop([(3, 0.5)], min_score=0.4, frame_index=50) == -1          # frame_index > 50 only
    assert lo.detect_loop([(3, 0.2)], min_score=0.4, frame_index=100) == -1         # below 0.6 * min_score
    # PoseGraph.cpp:212-218: the first result seeds best_index even if a later one is the only one above threshold
    assert lo.detect_loop([(3, 0.1), (9, 0.5)], min_score=0.4, frame_index=100) == 9


def test_search_by_brief_thresholds():
    w = np.stack([feat(0x00, 0x00), feat(0xFF, 0xFF)])
    old = np.stack([feat(0x01, 0x00), feat(0x00, 0x00), feat(0x00, 0x0F)])
    idx, dist, st = lo.search_by_brief(w, old)
    # window 1 (all ones): distances 255, 256, 8 + 31*4 = 132 -> none below 128 -> no match
    assert idx[0] == 1 and dist[0] == 0 and st[0] == 1
    assert idx[1] == -1 and st[1] == 0
    near = np.stack([feat(0x00, 0x00)])
    far = old[2:3].copy()          # distance 4 * 31 = 124 < 128 but >= 80: found, not accepted
    idx, dist, st = lo.search_by_brief(near, far)
    assert idx[0] == 0 and dist[0] == 124 and st[0] == 0


# ---------------------------------------------------------------- oracle pinned to the REFERENCE's own DBoW2 code
GOLDEN = ROOT / "tests" / "golden" / "dbow_golden.npz"


def test_bow_and_l1_score_equal_the_reference_dbow_goldens():
    """tests/golden/dbow_golden.npz comes from the reference's BowVector.cpp + ScoringObject.cpp compiled as they are
    (oracle/_ref/libref_dbow.so, tests/golden/make_dbow_golden.py): accumulation, normalisation and score bit for bit."""
    g = np.load(GOLDEN)
    n = int(g["n_seq"])
    bows = []
    for k in range(n):
        ids, vals = lo.bow_from_words(g[f"words_{k}"].tolist(), g[f"weights_{k}"].tolist())
        assert ids.tolist() == g[f"ids_{k}"].tolist()
        assert (vals == g[f"vals_{k}"]).all()
        bows.append((ids, vals))
    assert len(bows[0][0]) == 0                                    # the empty image
    for a in range(n):
        for b in range(n):
            assert lo.l1_score(*bows[a], *bows[b]) == g["scores"][a, b]
    assert all(g["scores"][k, k] == 1.0 or abs(g["scores"][k, k] - 1.0) < 1e-15 for k in range(1, n))


def test_database_query_scores_equal_reference_l1scoring_when_ref_is_built():
    """The database's inverted-file sum (TemplatedDatabase.h:587-646, restated) against L1Scoring::score of the reference
    on the same pair of vectors: the same terms in the same order."""
    so = ROOT / "oracle" / "_ref" / "libref_dbow.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libref_dbow.so not built (no reference tree here); the golden test above covers it")
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_dbow_golden as mg
    lib = mg.load_ref()
    v = lo.Vocabulary.random(4, 3, seed=3)
    frames, _ = make_keyframes(24, 6, per_image=60, seed=5)
    db = lo.Database(v)
    for f in frames:
        db.add(f)
    for q in (23, 11):
        qv = v.transform(frames[q])
        w, wt = zip(*[v.transform_feature(f) for f in frames[q]])
        rid, rval = mg.ref_bow(lib, w, wt)
        assert rid.tolist() == qv[0].tolist() and (rval == qv[1]).all()
        for e, s in db.query(frames[q], 0, -1):
            assert s == mg.ref_score(lib, qv, db.entries[e])


# ---------------------------------------------------------------- N > 1: entries sharded by id, one all_gather per query (gloo)
def _shard_query(voc, frames, q, max_id, rank, world):
    """What rank `rank` returns: its entries (global id % world == rank) scored, best four with GLOBAL ids - the contract of
    svin_loop_query on a context created with (rank, world)."""
    db = lo.Database(voc)
    gids = []
    for g, f in enumerate(frames):
        if g % world == rank:
            db.add(f)
            gids.append(g)
    local_max = -1 if max_id < 0 else sum(1 for g in gids if g < max_id)   # entries are added in id order
    return [(gids[e], s) for e, s in db.query(frames[q], 4, local_max)]


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from svin_b200.loop import merge_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    voc = lo.Vocabulary.random(4, 3, seed=3)
    frames, _ = make_keyframes(30, 6, per_image=60, seed=5)
    res = []
    for q, max_id in ((29, -1), (29, 20), (14, 9)):
        mine = torch.full((4, 2), -1.0, dtype=torch.float64)
        for k, (e, s) in enumerate(_shard_query(voc, frames, q, max_id, rank, world)):
            mine[k, 0], mine[k, 1] = e, s
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)               # the exchange step of the sharded query
        res.append(merge_shards([[(int(e), float(s)) for e, s in t.tolist() if e >= 0] for t in allr], 4))
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_query_merges_to_the_single_database_gloo():
    import multiprocessing as mp
    import socket
    voc = lo.Vocabulary.random(4, 3, seed=3)
    frames, _ = make_keyframes(30, 6, per_image=60, seed=5)
    db = lo.Database(voc)
    for f in frames:
        db.add(f)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for got, (qq, max_id) in zip(res, ((29, -1), (29, 20), (14, 9))):
        exp = db.query(frames[qq], 4, max_id)
        assert [e for e, _ in got] == [e for e, _ in exp]
        assert [s for _, s in got] == [float(s) for _, s in exp]


# ---------------------------------------------------------------- CUDA path against the oracle
def engine(v, **kw):
    from svin_b200.loop import LoopEngine
    return LoopEngine(v.first_child, v.num_children, v.descriptor, v.weight, v.word_id, **kw)


@pytest.mark.gpu
def test_words_and_bow_bit_exact():
    assert has_gpu()
    v = lo.Vocabulary.random(8, 3, seed=11)
    rng = np.random.default_rng(1)
    counts = [0, 1, 37, 500, 2048, 3]
    images = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in counts]
    images[5] = np.repeat(images[5][:1], 3, axis=0)        # the same feature three times: one word, summed weight
    with engine(v) as eng:
        got = eng.transform(images)
    for img, (ids, vals) in zip(images, got):
        eids, evals = v.transform(img)
        assert ids.tolist() == eids.tolist()
        assert (vals == evals).all()                        # f64, same operation order: bit-exact


@pytest.mark.gpu
def test_tiny_vocabulary_ties_on_device():
    assert has_gpu()
    v = tiny_vocabulary()
    half = np.zeros(32, np.uint8)
    half[:16] = 0xFF
    fs = [feat(0x03, 0x00), half, feat(0xF0, 0xFF), feat(0x0F, 0x00)]
    with engine(v) as eng:
        (ids, vals), = eng.transform([np.stack(fs)])
    eids, evals = v.transform(fs)
    assert ids.tolist() == eids.tolist() and (vals == evals).all()


@pytest.mark.gpu
@pytest.mark.parametrize("max_id", [-1, 40, 1, 0])
def test_query_bit_exact_small(max_id):
    assert has_gpu()
    v = lo.Vocabulary.random(6, 4, seed=2)
    frames, _ = make_keyframes(64, 9, per_image=120, seed=7)
    db = lo.Database(v)
    for f in frames:
        db.add(f)
    with engine(v) as eng:
        eng.add(frames[:30])
        eng.add(frames[30:])                                # incremental adds, as PoseGraph::addKeyFrame does
        assert eng.stats()["entries_total"] == 64
        for q in (63, 50, 10):
            got = eng.query(frames[q], 4, max_id)
            exp = db.query(frames[q], 4, max_id)
            assert [e for e, _ in got] == [e for e, _ in exp]
            assert [s for _, s in got] == [s for _, s in exp]
        # more results than candidates, and a query without features
        exp = db.query(frames[5], 64, 8)
        got = eng.query(frames[5], 64, 8)
        assert got == exp
        assert eng.query(np.zeros((0, 32), np.uint8), 4, -1) == []


@pytest.mark.gpu
def test_sharded_query_merges_to_the_single_database():
    """Entries sharded by id % world_size (one ctx per rank; here all on cuda:0), 4 results per rank, merged by the caller."""
    assert has_gpu()
    from svin_b200.loop import merge_shards
    v = lo.Vocabulary.random(6, 4, seed=2)
    frames, _ = make_keyframes(90, 10, per_image=100, seed=9)
    with engine(v) as single:
        single.add(frames)
        shards = [engine(v, rank=r, world=3) for r in range(3)]
        try:
            for s in shards:
                s.add(frames)
            assert sum(s.stats()["entries_local"] for s in shards) == 90
            for q in (89, 70, 33):
                for max_id in (-1, q - 20):
                    merged = merge_shards([s.query(frames[q], 4, max_id) for s in shards], 4)
                    assert merged == single.query(frames[q], 4, max_id)
        finally:
            for s in shards:
                s.close()


@pytest.mark.gpu
def test_brief_search_bit_exact():
    assert has_gpu()
    v = tiny_vocabulary()
    frames, _ = make_keyframes(2, 1, per_image=700, seed=4, flip_bits=30, fresh=0.3)
    with engine(v) as eng:
        for w, o in ((frames[0], frames[1]), (frames[0][:5], frames[1][:1]), (frames[0][:3], frames[1][:0])):
            idx, dist, st = eng.brief_search(w, o)
            eidx, edist, est = lo.search_by_brief(w, o)
            assert idx.tolist() == eidx.tolist()
            assert dist[eidx >= 0].tolist() == edist[eidx >= 0].tolist()
            assert st.tolist() == est.tolist()
            if len(o) == 700:
                assert 0 < st.sum() < len(st)       # both outcomes occur


@pytest.mark.gpu
def test_errors_are_reported_not_swallowed():
    assert has_gpu()
    from svin_b200 import capi
    from svin_b200.loop import LoopEngine
    v = tiny_vocabulary()
    with engine(v) as eng:
        with pytest.raises(capi.SvinError, match="features"):
            eng.transform([np.zeros((2049, 32), np.uint8)])
        with pytest.raises(capi.SvinError, match="max_results"):
            eng.query(np.zeros((3, 32), np.uint8), 0, -1)
    bad = v.first_child.copy()
    bad[1] = 0                                              # child index not after its parent: a cycle
    with pytest.raises(capi.SvinError, match="children"):
        LoopEngine(bad, v.num_children, v.descriptor, v.weight, v.word_id)


@pytest.mark.gpu
def test_configs4_size_5k_keyframes():
    """BASELINE configs[4]: 5k-keyframe database, k = 10 / L = 6 vocabulary (the shape of brief_k10L6.bin), 500 BRIEF
    descriptors per keyframe, db.query(bowVec, ret, 4, frame_index - 50) - against the vectorised oracle."""
    assert has_gpu()
    v = lo.Vocabulary.random(10, 6, seed=1)
    n = 5000
    frames, place = make_keyframes(n, 1200, per_image=500, seed=21, revisit_after=1700)
    db = lo.Database(v, fast=True)
    for f in frames:
        db.add(f)
    with engine(v) as eng:
        for s in range(0, n, 500):
            eng.add(frames[s:s + 500])
        got_bow = eng.transform(frames[4000:4040])
        for k, (ids, vals) in enumerate(got_bow):
            assert ids.tolist() == db.entries[4000 + k][0].tolist()
            assert (vals == db.entries[4000 + k][1]).all()
        loops = 0
        for q in range(n - 1, n - 400, -17):
            got = eng.query(frames[q], 4, q - 50)
            exp = db.query(frames[q], 4, q - 50)
            assert [e for e, _ in got] == [e for e, _ in exp]
            assert [s for _, s in got] == [s for _, s in exp]
            if got and place[got[0][0]] == place[q]:
                loops += 1
        assert loops >= 20           # the revisits are found: the data exercises the path, not an empty intersection
