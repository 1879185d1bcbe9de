"""CPU pinning of the DBoW2 transform oracle (oracle/orc_bow.cpp; TemplatedVocabulary.h:1127-1195,
:1218-1258, BowVector.cpp, FeatureVector.cpp, FORB.cpp) and of the kernels' source (csrc/bow_core.h run
single-threaded through bow_debug_host): an independent Python reading with dicts, bit-exact doubles
between the sorted-segment formulation and the std::map oracle, and the container properties."""
import numpy as np
import pytest

from orb_slam3_b200 import scenes

KEYS = ("bow_ids", "bow_vals", "fv_node_ids", "fv_ptr", "fv_idx")


def _descriptors(voc, n, seed):
    """Half random, half noisy copies of leaf descriptors (several features per word, some on stopped words)."""
    rng = np.random.default_rng(seed)
    k = voc._keep
    leaves = np.nonzero(k["child_ptr"][1:] == k["child_ptr"][:-1])[0]
    pick = rng.choice(leaves, n // 2)
    near = k["desc"][pick].copy()
    bits = np.unpackbits(near, axis=1)
    for r in range(len(bits)):
        bits[r, rng.choice(256, 3, replace=False)] ^= 1
    near = np.packbits(bits, axis=1)
    rnd = rng.integers(0, 256, (n - n // 2, 32), dtype=np.uint8)
    d = np.concatenate([near, rnd])
    return np.ascontiguousarray(d[rng.permutation(n)])


def _python_transform(voc, desc, levelsup):
    k = voc._keep
    cp, ci, nd, wt, wid = k["child_ptr"], k["child_ids"], k["desc"], k["weight"], k["word_id"]
    bits = np.unpackbits(nd, axis=1)
    v, fv = {}, {}
    for i, f in enumerate(desc):
        fb = np.unpackbits(f)
        node, lvl, nid = 0, 0, 0
        while cp[node] != cp[node + 1]:
            lvl += 1
            ch = ci[cp[node]:cp[node + 1]]
            d = (bits[ch] != fb).sum(1)
            node = int(ch[int(np.argmin(d))])              # argmin = first minimum
            if lvl == voc.L - levelsup:
                nid = node
        if wt[node] > 0:
            v[int(wid[node])] = v[int(wid[node])] + wt[node] if int(wid[node]) in v else wt[node]
            fv.setdefault(nid, []).append(i)
    norm = 0.0
    for key in sorted(v):
        norm += abs(v[key])
    ids = np.array(sorted(v), np.int32)
    vals = np.array([v[key] / norm for key in sorted(v)])
    return ids, vals, fv


@pytest.mark.parametrize("k,L,levelsup", [(6, 3, 1), (10, 4, 2), (4, 5, 4)])
def test_oracle_matches_python_dict_reading(oracle, k, L, levelsup):
    voc = scenes.synth_vocabulary(k, L, seed=L)
    desc = _descriptors(voc, 400, seed=k)
    r = oracle.bow_transform(voc, desc, levelsup)
    ids, vals, fv = _python_transform(voc, desc, levelsup)
    assert np.array_equal(r["bow_ids"], ids) and np.array_equal(r["bow_vals"], vals)
    assert list(r["fv_node_ids"]) == sorted(fv)
    for j, node in enumerate(r["fv_node_ids"]):
        assert list(r["fv_idx"][r["fv_ptr"][j]:r["fv_ptr"][j + 1]]) == fv[int(node)]
    assert r["used"] == sum(len(x) for x in fv.values()) < 400   # some features fell on stopped words
    assert len(ids) < r["used"]                                   # several features share a word


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (8, 3, 4, 777), (10, 5, 4, 1200), (3, 6, 3, 1), (3, 6, 3, 2), (10, 4, 2, 2048)])
def test_kernel_source_on_host_equals_oracle_bitwise(oracle, k, L, levelsup, n):
    from orb_slam3_b200 import bow
    voc = scenes.synth_vocabulary(k, L, seed=7)
    desc = _descriptors(voc, n, seed=n)
    ref = oracle.bow_transform(voc, desc, levelsup)
    got = bow.debug_host(voc, desc, levelsup)
    assert got["used"] == ref["used"]
    for key in KEYS:
        assert np.array_equal(got[key], ref[key]), key            # doubles compared bit for bit
    if got["used"]:
        assert abs(got["bow_vals"].sum() - 1.0) < 1e-12           # L1-normalised
    assert sorted(got["fv_idx"]) == sorted(set(got["fv_idx"])) and len(got["fv_idx"]) == got["used"]
    if levelsup >= L and got["used"]:
        assert list(got["fv_node_ids"]) == [0]                    # nid_level <= 0: everything under the root


def test_real_orb_descriptors_and_empty_input(oracle):
    from orb_slam3_b200 import bow
    from orb_slam3_b200.synth import synth_frame
    voc = scenes.synth_vocabulary(10, 4, seed=2)
    _, d, _ = oracle.OracleExtractor(1000).extract(synth_frame(480, 640, 3))
    ref, got = oracle.bow_transform(voc, d, 2), bow.debug_host(voc, d, 2)
    for key in KEYS:
        assert np.array_equal(got[key], ref[key]), key
    e = bow.debug_host(voc, np.zeros((0, 32), np.uint8), 2)
    assert e["used"] == 0 and len(e["bow_ids"]) == 0 and list(e["fv_ptr"]) == [0]
