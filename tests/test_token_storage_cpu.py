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


def _reference_encode_loop(run_folder, batches, token_base_size, token_dimensions, token_dtype, ragged_as_object):
    """Restatement of the reference's storage-filling loop (matchmaker/dense_retrieval.py:201-205 + :237-265 + :279-285)
    and of its `saveCompressed` (utils/utils.py:196-203), statement for statement, used ONLY to produce a folder that
    looks exactly like one the reference wrote (same file names, same four npz keys, same object layouts).
    `batches` = list of (seq_ids, output array [B, dim] or [B, L, dim])."""
    import os
    import zipfile
    token_base_number = 0
    token_base = np.memmap(os.path.join(run_folder, "token_reps_" + str(token_base_number) + ".npy"),
                           dtype=np.dtype(token_dtype), mode="w+", shape=(token_base_size, token_dimensions))
    current_ids = np.ndarray(shape=(token_base_size), dtype="int64")
    id_mapping, token_insert_index, storage, storage_filled_to_index = [], 0, [], []
    doc_infos, seq_ids = {}, []
    for ids, output in batches:
        for sample_i, seq_id in enumerate(ids):
            current_reps = output[sample_i]
            dim_count = len(current_reps.shape)
            if dim_count == 2:
                current_reps = current_reps[np.abs(current_reps).sum(-1) > 0, :]
            vec_count = 1 if dim_count == 1 else current_reps.shape[0]
            if token_insert_index + vec_count > token_base_size:
                storage.append(token_base[:token_insert_index])
                id_mapping.append(current_ids[:token_insert_index])
                current_ids = np.ndarray(shape=(token_base_size), dtype="int64")
                storage_filled_to_index.append(token_insert_index)
                token_base_number += 1
                token_insert_index = 0
                token_base = np.memmap(os.path.join(run_folder, "token_reps_" + str(token_base_number) + ".npy"),
                                       dtype=np.dtype(token_dtype), mode="w+", shape=(token_base_size, token_dimensions))
            start_index = token_insert_index
            token_insert_index = token_insert_index + vec_count
            token_base[start_index:token_insert_index] = current_reps
            current_ids[start_index:token_insert_index] = len(seq_ids)
            doc_infos[seq_id] = (token_base_number, start_index, token_insert_index)
            seq_ids.append(seq_id)
    storage.append(token_base[:token_insert_index])
    id_mapping.append(current_ids[:token_insert_index])
    storage_filled_to_index.append(token_insert_index)
    for s in storage:
        s.flush()
    if ragged_as_object:   # numpy < 1.24 (the reference's era) turned the ragged list into an object array by itself
        idm = np.empty(len(id_mapping), dtype=object)
        for i, x in enumerate(id_mapping):
            idm[i] = x
    else:
        idm = id_mapping
    with zipfile.ZipFile(os.path.join(run_folder, "doc_infos.npz"), mode="w", compression=zipfile.ZIP_STORED,
                         allowZip64=True) as zf:
        for k, v in dict(doc_infos=doc_infos, id_mapping=idm, seq_ids=seq_ids,
                         storage_filled_to_index=storage_filled_to_index).items():
            with zf.open(k + ".npy", "w", force_zip64=True) as buf:
                np.lib.format.write_array(buf, np.asanyarray(v), allow_pickle=True)
    return doc_infos, seq_ids


def test_reads_a_reference_written_folder_multi_block(tmp_path):
    """doc_infos.npz with ONLY the reference's four keys, several blocks (ragged id_mapping), multi-vector model:
    shapes and dtype come from the config arguments, as at dense_retrieval.py:299-300."""
    rng = np.random.default_rng(1)
    batches = []
    for b in range(5):
        out = rng.standard_normal((4, 6, 16)).astype(np.float32)
        out[:, 4:] = 0                      # padding token vectors (colbert doc_encode zeroes them)
        out[1, 2:] = 0
        batches.append(([f"d{b}_{i}" for i in range(4)], out))
    doc_infos, seq_ids = _reference_encode_loop(str(tmp_path), batches, 30, 16, "float16", ragged_as_object=True)
    z = np.load(str(tmp_path / "doc_infos.npz"), allow_pickle=True)
    assert sorted(z.files) == ["doc_infos", "id_mapping", "seq_ids", "storage_filled_to_index"]
    storage, id_mapping, got_ids, got_infos = load_token_storage(str(tmp_path), token_dim=16, token_block_size=30,
                                                                 token_dtype="float16")
    assert got_ids == seq_ids and got_infos == doc_infos and len(storage) >= 3
    for (ids, out) in batches:
        for i, sid in enumerate(ids):
            blk, lo, hi = got_infos[sid]
            exp = out[i][np.abs(out[i]).sum(-1) > 0].astype(np.float16)
            assert np.array_equal(np.asarray(storage[blk][lo:hi]), exp)
            assert all(got_ids[j] == sid for j in id_mapping[blk][lo:hi])
    # without the config values there is nothing to read the block shape from: must raise, not guess
    import pytest
    with pytest.raises(ValueError):
        load_token_storage(str(tmp_path))


def test_reads_a_reference_written_folder_single_block(tmp_path):
    """One block, single-vector model (BERT_DOT): id_mapping is saved as a 2-d int64 array by numpy.asanyarray."""
    rng = np.random.default_rng(2)
    out = rng.standard_normal((9, 8)).astype(np.float32)
    doc_infos, seq_ids = _reference_encode_loop(str(tmp_path), [([f"p{i}" for i in range(9)], out)], 100, 8, "float32",
                                                ragged_as_object=False)
    storage, id_mapping, got_ids, got_infos = load_token_storage(str(tmp_path), 8, 100, "float32")
    assert len(storage) == 1 and storage[0].shape == (9, 8) and np.array_equal(np.asarray(storage[0]), out)
    assert id_mapping[0].tolist() == list(range(9)) and got_ids == seq_ids and got_infos == doc_infos


def test_our_writer_emits_exactly_the_reference_keys(tmp_path):
    w = TokenStorageWriter(str(tmp_path), token_dim=4, token_block_size=5)
    for i in range(4):
        w.add(f"x{i}", np.ones((3, 4), dtype=np.float16) * (i + 1))
    w.close()
    z = np.load(str(tmp_path / "doc_infos.npz"), allow_pickle=True)
    assert sorted(z.files) == ["doc_infos", "id_mapping", "seq_ids", "storage_filled_to_index"]
    # the reference's own restore statements (dense_retrieval.py:292-295) work on it
    doc_infos = z.get("doc_infos")[()]
    id_mapping = z.get("id_mapping")[()]
    filled = z.get("storage_filled_to_index")[()]
    assert doc_infos["x3"] == (3, 0, 3) and len(id_mapping) == 4 and list(filled) == [3, 3, 3, 3]
    m = np.memmap(str(tmp_path / "token_reps_2.npy"), dtype=np.float16, mode="r", shape=(5, 4))[:filled[2]]
    assert (np.asarray(m) == 3).all()
