"""On-disk formats either side of the hot path.

* descriptor shards: the search operands of an :class:`~cirtorch_b200.search.Index` (K-major bf16 rows, optionally the
  fp32 rows kept for exact re-scoring) as one ``torch.save`` file per shard, loadable straight into HBM;
* the reference's dataset pickles, read with the reference's own field names:
  ``gnd_<dataset>.pkl`` (cirtorch/datasets/globalFeatures/oxford_paris.py:9-34) and the retrieval-SfM training
  pickle (``cids`` / ``cluster`` / ``qidxs`` / ``pidxs``, tuples_dataset.py:60-92);
* ``ret_head`` weights inside a reference snapshot (cirtorch/utils/snapshot.py:6-17,41-75).
"""
from __future__ import annotations

import pickle
from os import path

import torch

from .search import Index

FORMAT = "cirtorch_b200.index.v1"
TEST_DATASETS = ["oxford5k", "paris6k", "roxford5k", "rparis6k"]


def save_index(file, index: Index, extra: dict = None):
    """Write one database shard (packed bf16 rows [+ fp32 rows, labels]) with its geometry."""
    data = {"format": FORMAT, "mode": index.mode, "N": index.N, "D": index.D, "row_offset": index.row_offset,
            "packed": index.packed.cpu(), "rows32": None if index.rows32 is None else index.rows32.cpu(),
            "labels": None if index.labels is None else index.labels.cpu(), "extra": extra or {}}
    torch.save(data, file)


def load_index(file, device="cuda", keep_fp32=True) -> Index:
    """Load a shard written by :func:`save_index` into device memory without re-packing."""
    data = torch.load(file, map_location="cpu")
    if data.get("format") != FORMAT:
        raise ValueError("%s is not a %s file" % (file, FORMAT))
    index = Index.__new__(Index)
    index.mode, index.N, index.D, index.row_offset = data["mode"], data["N"], data["D"], data["row_offset"]
    index.packed = data["packed"].to(device)
    index.rows32 = data["rows32"].to(device) if (keep_fp32 and data["rows32"] is not None) else None
    index.labels = None if data["labels"] is None else data["labels"].to(device)
    return index


def load_test_dataset(root_dir, name):
    """``ParisOxfordTestDataset`` (oxford_paris.py:9-34): the gnd pickle plus the derived path / count fields."""
    if name not in TEST_DATASETS:
        raise ValueError("Unknown dataset: {}!".format(name))
    pkl_path = path.join(root_dir, "gnd_{}.pkl".format(name))
    with open(pkl_path, "rb") as f:
        db = pickle.load(f)
    db["pkl_path"] = pkl_path
    db["data_path"] = path.join(root_dir)
    db["images_path"] = path.join(db["data_path"], "jpg")
    db["_ext"] = ".jpg"
    db["n_img"] = len(db["imlist"])
    db["n_query"] = len(db["qimlist"])
    db["img_names"] = [path.join(db["images_path"], item + db["_ext"]) for item in db["imlist"]]
    db["query_names"] = [path.join(db["images_path"], item + db["_ext"]) for item in db["qimlist"]]
    db["query_bbx"] = [db["gnd"][item]["bbx"] for item in range(db["n_query"])]
    db["dataset"] = name
    return db


def load_training_db(db_fn, mode="train"):
    """The retrieval-SfM / gl training pickle (tuples_dataset.py:60-92): returns the fields mining needs."""
    if mode not in ("train", "val"):
        raise RuntimeError("Mode should be either train or val, passed as string")
    with open(db_fn, "rb") as f:
        db = pickle.load(f)[mode]
    return {"cids": db["cids"], "cluster": db["cluster"], "qidxs": db["qidxs"], "pidxs": db["pidxs"]}


def miner_from_training_db(db, nnum=5, qsize=2000, poolsize=20000, images=None, transform=None):
    from .mining import TuplesMiner
    return TuplesMiner(db["cluster"], db["qidxs"], db["pidxs"], neg_num=nnum, query_size=qsize, pool_size=poolsize,
                       images=images, transform=transform)


def load_ret_head(snapshot_file, head, strict_shapes=False):
    """Load ``state_dict['ret_head']`` of a reference snapshot (snapshot.py:6-17) into a globalHead.  Like the
    reference's ``_load_pretraining_dict`` (:54-75), parameters whose shapes differ are skipped unless ``strict_shapes``."""
    snapshot = torch.load(snapshot_file, map_location="cpu")
    state = dict(snapshot["state_dict"]["ret_head"])
    model_sd = head.state_dict()
    for k, v in model_sd.items():
        if k in state and v.shape != state[k].shape:
            if strict_shapes:
                raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(v.shape), tuple(state[k].shape)))
            del state[k]
    head.load_state_dict(state, False)
    return snapshot.get("training_meta", {})
