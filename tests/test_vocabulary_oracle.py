"""CPU: the oracle's restatement of DBoW2 TemplatedVocabulary::transform against the reference's own DBoW2, compiled
from /root/reference/Thirdparty/DBoW2 (oracle/_ref/libref_dbow.so) and fed the same vocabulary through the fork's
text loader -- BowVector (ids and doubles, bit for bit) and FeatureVector identical."""
import os

import numpy as np
import pytest

import oracle
from voc_cases import features_near_leaves, make_vocabulary, write_text

HAVE_REF = os.path.exists(oracle._REF_DBOW) or os.path.exists("/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h")


@pytest.mark.skipif(not HAVE_REF, reason="reference DBoW2 build not available")
@pytest.mark.parametrize("k,L,weighting,scoring,levelsup", [(10, 4, 0, 0, 2), (10, 3, 0, 0, 4), (5, 5, 1, 5, 3), (8, 3, 2, 1, 1),
                                                            (6, 4, 3, 0, 2), (10, 4, 0, 1, 4), (4, 6, 1, 0, 4)])
def test_oracle_transform_equals_reference_dbow2(tmp_path, k, L, weighting, scoring, levelsup):
    rng = np.random.default_rng(1000 * k + 100 * L + 10 * weighting + scoring)
    v = make_vocabulary(rng, k, L, weighting, scoring)
    path = str(tmp_path / "voc.txt")
    write_text(v, path)
    ref = oracle.RefVocabulary(path)
    assert ref.size() == int(v.leaf.sum())
    feats = features_near_leaves(rng, v, 1500)
    a = ref.transform(feats, levelsup)
    b = oracle.voc_transform(v, feats, levelsup)
    assert (a["bow_ids"] == b["bow_ids"]).all()
    assert a["bow_values"].tobytes() == b["bow_values"].tobytes()   # doubles, bit for bit
    assert a["fv"] == b["fv"]
    assert len(a["bow_ids"]) > 50 and len(a["fv"]) >= (1 if L - levelsup <= 0 else 3)


def test_oracle_transform_properties():
    rng = np.random.default_rng(5)
    v = make_vocabulary(rng, 10, 4)
    feats = features_near_leaves(rng, v, 800)
    r = oracle.voc_transform(v, feats, 2)
    assert abs(np.abs(r["bow_values"]).sum() - 1.0) < 1e-12       # L1 normalised
    assert (np.diff(r["bow_ids"]) > 0).all()
    listed = sorted(i for idx in r["fv"].values() for i in idx)
    keep = v.node_weight[[int(np.nonzero((v.node_word == w) & v.leaf)[0][0]) for w in r["words"]]] > 0
    assert listed == list(np.nonzero(keep)[0])                     # every non-stopped feature exactly once
    # a feature that equals a leaf descriptor lands on that leaf (or an equal-distance earlier sibling chain)
    leaf_ids = np.nonzero(v.leaf)[0][:50]
    r2 = oracle.voc_transform(v, v.node_desc[leaf_ids], 2)
    d = [int(np.unpackbits(v.node_desc[l] ^ v.node_desc[np.nonzero((v.node_word == w) & v.leaf)[0][0]]).sum()) for l, w in zip(leaf_ids, r2["words"])]
    assert sum(x == 0 for x in d) >= 40
