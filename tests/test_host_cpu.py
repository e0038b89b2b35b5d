"""CPU-only checks: the C-ABI library loads and exports every symbol of include/cir_b200.h, the
argument validation of the host layer, and that nothing falls back to the CPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cir_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from cirtorch_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "binding missing for " + name
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.cir_version() >= 100


def test_no_cpu_fallback():
    from cirtorch_b200._lib import CirError
    from cirtorch_b200.modules.pools import GeM
    from cirtorch_b200.modules.normalizations import L2N
    from cirtorch_b200 import search as S
    x = torch.rand(2, 8, 4, 4)
    with pytest.raises(CirError):
        GeM()(x)
    with pytest.raises(CirError):
        L2N()(x)
    with pytest.raises(CirError):
        S.search_topk(torch.rand(64, 3), torch.rand(64, 10), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "image-retrieval-for-image-based-localization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_registries_and_signatures_mirror_reference():
    from cirtorch_b200.modules.pools import POOLING_LAYERS, GeM, GeMmp
    from cirtorch_b200.modules.normalizations import NORMALIZATION_LAYERS
    from cirtorch_b200.modules.heads.global_head import globalHead
    from cirtorch_b200.layers.pooling import GeM as GeM2
    import inspect
    assert {"MAC", "SPoC", "GeM", "GeMmp"} <= set(POOLING_LAYERS) and "L2N" in NORMALIZATION_LAYERS
    assert GeM is GeM2
    assert list(inspect.signature(GeM.__init__).parameters)[1:] == ["p", "eps"]
    assert list(inspect.signature(GeMmp.__init__).parameters)[1:] == ["p", "mp", "eps"]
    assert list(inspect.signature(globalHead.__init__).parameters)[1:] == ["pooling", "normal", "dim", "norm_act"]
    fwd = inspect.signature(globalHead.forward).parameters
    positional = [n for n, a in fwd.items() if a.kind == a.POSITIONAL_OR_KEYWORD][1:]
    assert positional == ["x", "do_whitening"]                      # the reference's call signature
    assert all(a.default is not a.empty for n, a in fwd.items() if a.kind == a.KEYWORD_ONLY)   # extensions are optional
    from cirtorch_b200.modules.pools import RMAC, Rpool
    assert POOLING_LAYERS["RMAC"] is RMAC and POOLING_LAYERS["ROIpool"] is Rpool
    assert list(inspect.signature(RMAC.__init__).parameters)[1:] == ["L", "eps"]
    assert list(inspect.signature(Rpool.__init__).parameters)[1:] == ["rpool", "whiten", "L", "eps"]
    head = globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=32)
    assert set(head.state_dict()) == {"pool.p", "whiten.weight", "whiten.bias"}
    assert float(head.whiten.bias.detach().abs().max()) == 0.0 and float(head.whiten.weight.detach().std()) < 0.05


def test_shard_bounds_partition():
    from cirtorch_b200.parallel import shard_bounds
    for n in (0, 1, 7, 1_000_000, 4993):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_store_formats_roundtrip(tmp_path):
    """N3: dataset pickles with the reference's field names, snapshot ret_head loading (CPU only)."""
    import pickle
    import numpy as np
    from cirtorch_b200 import store
    from cirtorch_b200.modules.heads.global_head import globalHead
    gnd = {"imlist": ["a", "b", "c"], "qimlist": ["q0"], "gnd": [{"bbx": [0, 0, 10, 10], "easy": [0], "hard": [1], "junk": [2]}]}
    with open(tmp_path / "gnd_roxford5k.pkl", "wb") as f:
        pickle.dump(gnd, f)
    db = store.load_test_dataset(str(tmp_path), "roxford5k")
    assert db["n_img"] == 3 and db["n_query"] == 1 and db["query_bbx"] == [[0, 0, 10, 10]]
    assert db["img_names"][0].endswith("jpg/a.jpg") and db["dataset"] == "roxford5k"
    with pytest.raises(ValueError):
        store.load_test_dataset(str(tmp_path), "nope")
    train = {"train": {"cids": ["x"] * 6, "cluster": [0, 0, 1, 1, 2, 2], "qidxs": [0, 2], "pidxs": [1, 3]}, "val": {}}
    with open(tmp_path / "sfm.pkl", "wb") as f:
        pickle.dump(train, f)
    tdb = store.load_training_db(str(tmp_path / "sfm.pkl"), "train")
    miner = store.miner_from_training_db(tdb, nnum=1, qsize=2, poolsize=6)
    assert miner.query_size == 2 and miner.pool_size == 6 and list(miner.clusters) == [0, 0, 1, 1, 2, 2]
    assert miner.neg_num == 1 and miner.negative_indices is None and miner.nidxs is None
    head = globalHead(pooling={"name": "GeM", "params": {"p": 3, "eps": 1e-6}}, normal={"name": "L2N", "params": {}}, dim=8)
    sd = {k: v.clone() + 1.0 for k, v in head.state_dict().items()}
    sd["whiten.bias"] = torch.zeros(3)                     # wrong shape: skipped like _load_pretraining_dict
    torch.save({"config": "", "state_dict": {"ret_head": sd, "body": {}}, "training_meta": {"epoch": 7}}, tmp_path / "snap.pth")
    meta = store.load_ret_head(str(tmp_path / "snap.pth"), head)
    assert meta["epoch"] == 7 and float(head.pool.p.detach()) == 4.0 and float(head.whiten.bias.detach().abs().max()) == 0.0


def test_rmac_region_grid_matches_fixture(golden):
    """Host logic of the regional pooling: the product's region grid == the grid recovered from the reference."""
    from cirtorch_b200 import functional as LF
    grid = golden("regional")["grid"].tolist()
    for (H, W, L) in sorted({tuple(r[:3]) for r in grid}):
        want = [tuple(r[3:]) for r in grid if tuple(r[:3]) == (H, W, L)]
        assert [(0, 0, H, W)] + LF.rmac_regions(H, W, L) == want, (H, W, L)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) must print ONE JSON line with the
    contract keys, without a GPU.  One step of the oracle port of globalHead.forward on the full 64 x 2048 x 32 x 32 batch."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "descriptors/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert key in line


def test_search_planner_invariants():
    """cir_search_plan is pure host logic: every database tile belongs to exactly one split, the unit count and the
    padding follow from the tiles, the list capacity leaves room beyond k, and the threshold sample obeys its rule."""
    import ctypes as C
    from cirtorch_b200 import _lib
    lib = _lib.load()
    out = (C.c_int32 * 8)()
    for num_sms in (148, 132, 16):
        for Q in (1, 70, 128, 129, 1000, 2049, 10_000):
            for N in (1, 255, 4993, 8192, 20_000, 65_536, 125_000, 1_000_000, 3_000_001):
                for k in (1, 10, 100, 512):
                    assert lib.cir_search_plan(Q, N, k, num_sms, C.cast(out, C.c_void_p)) == 0
                    mt, nt, S, tps, units, Qpad, cap, n0 = list(out)
                    assert mt == -(-Q // 128) and nt == -(-N // 256) and Qpad == mt * 128
                    assert S >= 1 and tps >= 1 and S * tps >= nt and (S - 1) * tps < nt and units == mt * S
                    assert cap & (cap - 1) == 0 and cap >= k + 96 and cap <= 1024
                    if n0:
                        assert N >= 8192 and n0 % 256 == 0 and 1024 <= n0 <= 32768 and 2 * n0 <= N
                        assert n0 // 8 >= 2 * k                      # at least 2 k groups of 8 rows
                        assert not (Q <= 128 and N < 65_536)
                    else:
                        assert N < 8192 or (Q <= 128 and N < 65_536) or k > 64     # small, single-tile, or k too large for N
    assert lib.cir_search_plan(0, 10, 1, 148, C.cast(out, C.c_void_p)) != 0
    assert b"cir_search_plan" in lib.cir_last_error()


def test_entry_points_validate_before_touching_the_device():
    """Error behaviour of the C ABI without a GPU: null pointers / bad shapes come back as CIR_ERR_INVALID_ARG (-1) with a
    message naming the entry point -- the checks run before any CUDA call, so no compute is attempted here."""
    import ctypes as C
    from cirtorch_b200 import _lib
    lib = _lib.load()
    assert lib.cir_version() >= 100 and lib.cir_launch_count(0) >= 0
    need = C.c_size_t(0)
    assert lib.cir_tail_workspace_bytes(0, 2048, 2048, C.byref(need)) == -1
    assert lib.cir_tail_workspace_bytes(64, 2048, 2048, C.byref(need)) == 0 and need.value > 64 * 2048 * 4
    calls = {
        "cir_tail_fwd": lambda: lib.cir_tail_fwd(None, 1, 8, 4, 4, None, 0, 1e-6, 1e-6, 0, None, None, 8, None, 8, None, None, 0, 0, None),
        "cir_tail_fwd_train": lambda: lib.cir_tail_fwd_train(None, 1, 8, 4, 4, None, 0, 1e-6, 1e-6, 0, None, None, 8, None, 8, None,
                                                             C.c_void_p(256), None, 0, _lib.CIR_TAIL_NO_WHITEN, None),
        "cir_search_topk_exchange_merge": lambda: lib.cir_search_topk_exchange_merge(None, 70, None, 10, 64, 5, 0, None, 1, 0, 1, None,
                                                                                     None, None, 0, 0, None),
        "cir_gem_bwd": lambda: lib.cir_gem_bwd(None, 1, 8, 4, 4, None, 0, 1e-6, None, None, None, None, None),
        "cir_region_pool": lambda: lib.cir_region_pool(None, 1, 8, 4, 4, None, 1, None, 0, 1e-6, 0, None, None),
        "cir_l2n_rows": lambda: lib.cir_l2n_rows(None, 1, 8, 8, 1e-6, None, 8, None),
        "cir_pack_bf16": lambda: lib.cir_pack_bf16(None, 1, 8, 8, None, 64, 1, 0, None),
        "cir_search_workspace_bytes": lambda: lib.cir_search_workspace_bytes(0, 10, 64, 5, C.byref(need)),
        "cir_topk_merge": lambda: lib.cir_topk_merge(None, None, 2, 4, 5, 0, None, None, 5, None),
        "cir_rescore_topk": lambda: lib.cir_rescore_topk(None, 1, None, 10, 8, None, 4, 0, None, None, 2, None),
        "cir_qe_aggregate": lambda: lib.cir_qe_aggregate(None, 1, None, 10, 8, None, None, 4, 4, 2, 3.0, -1, 1e-6, None, None),
        "cir_mine_filter": lambda: lib.cir_mine_filter(None, 1, 4, None, 10, None, 2, None, None, 8, None, None, None, None, None, 0.0, None, None),
        "cir_eval_ap": lambda: lib.cir_eval_ap(None, 1, 4, 4, None, None, None, None, None, 0, None, None, None),
    }
    for name, call in calls.items():
        assert call() == -1, name
        assert name.encode() in lib.cir_last_error(), (name, lib.cir_last_error())


def test_ragged_batch_containers():
    """PackedSequence / pad_packed_images / pack_padded_images (host containers either side of the path): own invariants,
    and -- where the reference tree is present -- the same results as cirtorch/utils/parallel/packed_sequence.py and
    cirtorch/utils/sequence.py on random ragged inputs."""
    import os
    import sys
    import pytest
    import torch
    from cirtorch_b200.utils.sequence import PackedSequence, pad_packed_images, pack_padded_images
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randn(3, h, w, generator=g) for (h, w) in ((5, 7), (9, 4), (6, 6))]
    seq = PackedSequence(imgs)
    assert len(seq) == 3 and seq[1] is imgs[1] and not seq.all_none and seq.dtype == torch.float32
    assert isinstance(seq[0:2], PackedSequence) and len(seq[0:2]) == 2 and len(seq + seq) == 6
    padded, sizes = pad_packed_images(seq, pad_value=-1.0, snap_size_to=4)
    assert tuple(padded.shape) == (3, 3, 12, 8) and [tuple(s) for s in sizes] == [(5, 7), (9, 4), (6, 6)]
    assert float(padded[0, :, 5:, :].max()) == -1.0 and torch.equal(padded[1, :, :9, :4], imgs[1])
    back = pack_padded_images(padded, sizes)
    assert all(torch.equal(a, b) for a, b in zip(back, imgs))
    labels = PackedSequence([torch.tensor([-1., 1., 0.]), None, torch.tensor([-1., 1., 0.])])
    cat, idx = labels.contiguous
    assert cat.tolist() == [-1, 1, 0, -1, 1, 0] and idx.tolist() == [0, 0, 0, 2, 2, 2]
    assert PackedSequence([None, None]).contiguous == (None, None) and PackedSequence([None]).all_none
    with pytest.raises(ValueError):
        PackedSequence(imgs).contiguous                      # different spatial sizes: no contiguous view
    with pytest.raises(TypeError):
        PackedSequence([torch.zeros(1), torch.zeros(1, dtype=torch.float64)])
    with pytest.raises(ValueError):
        pad_packed_images(PackedSequence([None]))
    if os.path.isdir("/root/reference/cirtorch"):
        sys.path.insert(0, "/root/reference")
        from cirtorch.utils.parallel.packed_sequence import PackedSequence as RefSeq
        from cirtorch.utils.sequence import pad_packed_images as ref_pad
        for with_none in (False, True):
            items = list(imgs) + ([None] if with_none else [])
            a, sa = pad_packed_images(PackedSequence(items), pad_value=0.5, snap_size_to=None)
            b, sb = ref_pad(RefSeq(items), pad_value=0.5, snap_size_to=None)
            assert torch.equal(a, b) and [tuple(x) for x in sa] == [tuple(x) for x in sb]
        ra, rb = RefSeq([torch.ones(2, 3), None, torch.zeros(1, 3)]).contiguous
        ma, mb = PackedSequence([torch.ones(2, 3), None, torch.zeros(1, 3)]).contiguous
        assert torch.equal(ra, ma) and torch.equal(rb, mb)
