"""The oracle's DBoW2 transform (orc_bow.cpp = what Frame::ComputeBoW computes) against THE
REFERENCE'S OWN OBJECT CODE: the vendored Thirdparty/DBoW2 sources compiled unmodified (oracle/_ref/libref_bow.so,
oracle/Makefile), the vocabulary loaded by the reference's own loadFromTextFile from an ORBvoc.txt-format file, the
transform run by TemplatedVocabulary<FORB::TDescriptor, FORB>::transform.  Everything here is integer / exact-order
floating point, so the comparison is bit for bit: word ids, the L1-normalised double weights, node ids and feature
lists.  Built where /root/reference exists, shipped prebuilt to the GPU box; skipped when absent."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as R
    if not R.bow_available():
        pytest.skip("oracle/_ref/libref_bow.so is not built and the reference tree is absent")
    return R


def _arrays(v):
    a = v._keep
    return a["child_ptr"], a["child_ids"], a["desc"], a["weight"], a["word_id"]


@pytest.mark.parametrize("k,L,seed,irregular,levelsup", [(10, 4, 2, True, 4), (10, 4, 3, False, 4), (6, 5, 4, True, 4), (9, 3, 5, True, 2),
                                                         (10, 6, 6, True, 4)])
def test_transform_equals_the_reference_object_code(oracle, ref, tmp_path, k, L, seed, irregular, levelsup):
    if L == 6:
        k = 4   # a six-level tree like ORBvoc.txt's, kept small
    voc = scenes.synth_vocabulary(k, L, seed=seed, irregular=irregular)
    child_ptr, child_ids, desc, weight, word_id = _arrays(voc)
    path = str(tmp_path / "voc.txt")
    ref.write_vocabulary_text(path, k, L, child_ptr, child_ids, desc, weight)
    rv = ref.RefVocabulary(path)
    assert rv.size() == int((word_id >= 0).sum())          # words = leaves, in node order
    rng = np.random.default_rng(seed)
    for n in (1, 37, 1500):
        # features near vocabulary nodes (real descents with ties) and pure noise
        base = desc[rng.integers(1, len(desc), n)]
        bits = np.unpackbits(base, axis=1)
        flip = rng.random(bits.shape) < 0.08
        d = np.packbits(bits ^ flip, axis=1)
        d[::5] = rng.integers(0, 256, (len(d[::5]), 32), dtype=np.uint8)
        got = oracle.bow_transform(voc, d, levelsup)
        exp = rv.transform(d, levelsup)
        assert np.array_equal(got["bow_ids"], exp["bow_ids"])
        assert np.array_equal(got["bow_vals"], exp["bow_vals"]), np.abs(got["bow_vals"] - exp["bow_vals"]).max()   # identical doubles
        assert np.array_equal(got["fv_node_ids"], exp["fv_node_ids"])
        assert np.array_equal(got["fv_ptr"], exp["fv_ptr"]) and np.array_equal(got["fv_idx"], exp["fv_idx"])
