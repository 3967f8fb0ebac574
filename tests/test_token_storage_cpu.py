"""Round trip of the reference's dense-retrieval storage layout (token_reps_N.npy blocks + doc_infos.npz)."""
import numpy as np

from matchmaker_b200.retrieval.token_storage import TokenStorageWriter, load_token_storage


def test_roundtrip_single_and_multi_vector(tmp_path):
    rng = np.random.default_rng(0)
    w = TokenStorageWriter(str(tmp_path), token_dim=8, token_block_size=10)
    vecs = {}
    for i in range(7):
        if i % 2:
            v = rng.standard_normal(8).astype(np.float16)
        else:
            v = rng.standard_normal((4, 8)).astype(np.float16)
            v[1] = 0  # an all-zero (padding) row must be stripped
        w.add(f"doc{i}", v)
        vecs[f"doc{i}"] = v
    w.close()
    storage, id_mapping, seq_ids, doc_infos = load_token_storage(str(tmp_path))
    assert seq_ids == [f"doc{i}" for i in range(7)]
    assert len(storage) >= 2 and sum(len(s) for s in storage) == sum(len(m) for m in id_mapping)
    for name, v in vecs.items():
        b, lo, hi = doc_infos[name]
        got = np.asarray(storage[b][lo:hi])
        exp = v[np.newaxis, :] if v.ndim == 1 else v[np.abs(v).sum(-1) != 0]
        assert np.array_equal(got, exp)
        assert all(seq_ids[j] == name for j in id_mapping[b][lo:hi])
