"""GPU: DBoW2 vocabulary transform through the C ABI (tree descent on the device) against the oracle restatement and
against the reference's own DBoW2 compiled from source: BowVector doubles bit for bit, FeatureVector identical."""
import os

import numpy as np
import pytest

import oracle
from orb_slam_system_b200 import ORBextractor, ORBVocabulary
from voc_cases import features_near_leaves, make_vocabulary, write_text

pytestmark = pytest.mark.gpu


def _product(v):
    return ORBVocabulary(v.child_off, v.children, v.node_desc, v.node_weight, v.node_word, v.k, v.L, v.weighting, v.scoring)


@pytest.mark.parametrize("k,L,weighting,scoring,levelsup,early", [(10, 4, 0, 0, 2, 0.0), (10, 3, 0, 0, 4, 0.0), (5, 5, 1, 5, 3, 0.0),
                                                                  (8, 3, 2, 1, 1, 0.0), (6, 4, 3, 0, 2, 0.0), (20, 3, 0, 0, 1, 0.0),
                                                                  (10, 5, 0, 0, 4, 0.15)])
def test_transform_matches_oracle(k, L, weighting, scoring, levelsup, early):
    rng = np.random.default_rng(1000 * k + 100 * L + 10 * weighting + scoring)
    v = make_vocabulary(rng, k, L, weighting, scoring, early_leaf=early)
    feats = features_near_leaves(rng, v, 3000)
    voc = _product(v)
    got = voc.transform(feats, levelsup)
    want = oracle.voc_transform(v, feats, levelsup)
    assert (got["words"] == want["words"]).all() and (got["nodes"] == want["nodes"]).all()
    assert (got["bow_ids"] == want["bow_ids"]).all()
    assert got["bow_values"].tobytes() == want["bow_values"].tobytes()
    assert got["fv"] == want["fv"]
    voc.close()


@pytest.mark.skipif(not os.path.exists(oracle._REF_DBOW), reason="reference DBoW2 build not shipped")
def test_transform_matches_reference_dbow2_on_extracted_descriptors(tmp_path):
    # Frame::ComputeBoW: the descriptors of one extracted frame through transform(..., 4), ORBvoc-like k = 10
    rng = np.random.default_rng(77)
    img = oracle.synth_frame(480, 640, frame=3)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    _, desc = ex(img)
    v = make_vocabulary(rng, 10, 5)
    # seed some leaves with real descriptors so that words repeat
    leaves = np.nonzero(v.leaf)[0]
    v.node_desc[leaves[: len(desc)]] = desc[: len(leaves)]
    path = str(tmp_path / "voc.txt")
    write_text(v, path)
    ref = oracle.RefVocabulary(path)
    voc = ORBVocabulary.loadFromTextFile(path)
    assert voc.size() == ref.size()
    got = voc.transform(desc, 4)
    want = ref.transform(desc, 4)
    assert (got["bow_ids"] == want["bow_ids"]).all()
    assert got["bow_values"].tobytes() == want["bow_values"].tobytes()
    assert got["fv"] == want["fv"]
    ex.close()
    voc.close()


def test_transform_edge_cases():
    rng = np.random.default_rng(3)
    v = make_vocabulary(rng, 4, 2)
    voc = _product(v)
    r = voc.transform(np.zeros((0, 32), np.uint8))
    assert len(r["bow_ids"]) == 0 and r["fv"] == {}
    one = voc.transform(v.node_desc[5:6], 1)
    o = oracle.voc_transform(v, v.node_desc[5:6], 1)
    assert (one["bow_ids"] == o["bow_ids"]).all() and one["fv"] == o["fv"]
    voc.close()
