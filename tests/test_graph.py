"""Host-side graph path: generator vs the numpy oracle, reference file format, partition."""
import os

import numpy as np
import pytest

import hpc_b200 as H
from oracle import graph_oracle as G
from oracle import plan_oracle as P
import refshim


def test_generator_matches_oracle_small():
    for shape, seed in [(H.GRAPH_SHAPES["c0"], 123), ((300, 3000, 250, 2, 100000, 400000, 16), 5),
                        ((64, 64 * 40, 60, 2, 0, 500000, 4), 7), ((50, 49, 49, 4, 0, 0, 1), 1)]:
        ptr, idx = H.gen_graph(*shape, seed=seed)
        optr, oidx = G.gen_graph(*shape, seed)
        assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx), shape


def test_generator_named_shapes_degrees_and_sample_rows():
    for name in ("arxiv",):
        shape = H.GRAPH_SHAPES[name]
        ptr, idx = H.gen_named_graph(name)
        deg = np.diff(ptr)
        assert len(deg) == shape[0] and ptr[-1] == shape[1] and deg.max() == shape[2]
        assert np.array_equal(deg, G.gen_degrees(*shape[:5], 123))
        rows = [0, 17, 1000, int(np.argmax(deg)), shape[0] - 1]
        _, orows = G.gen_graph(*shape, 123, rows=rows)
        for r in rows:
            assert np.array_equal(idx[ptr[r]:ptr[r + 1]], orows[r]), r
        # CSR invariants: ascending, unique, in range
        starts = np.zeros(len(idx), bool)
        starts[ptr[:-1][deg > 0]] = True
        assert np.all((np.diff(idx.astype(np.int64)) > 0) | starts[1:])
        assert idx.min() >= 0 and idx.max() < shape[0]


def test_generator_full_size_statistics():
    """reddit / products shapes: exact (rows, nnz, max row nnz) of BASELINE.json + phase_2.log."""
    for name in ("reddit", "products"):
        m, nnz, mx = H.GRAPH_SHAPES[name][:3]
        deg = H.gen_degrees(*H.GRAPH_SHAPES[name][:5])
        assert len(deg) == m and int(deg.sum(dtype=np.int64)) == nnz and int(deg.max()) == mx


def test_generator_rejects_bad_arguments():
    with pytest.raises(H.SpmmB200Error):
        H.gen_graph(10, 5, 20)          # max_deg > num_v
    with pytest.raises(H.SpmmB200Error):
        H.gen_graph(10, 1000, 5)        # nnz unreachable


def test_graph_files_roundtrip(tmp_path):
    """PA4/handout/src/data.cu:3-66: text .graph + .config, dumps written on first read."""
    ptr, idx = H.gen_graph(200, 1500, 100, seed=3)
    d = str(tmp_path)
    H.write_graph(d, "g1", ptr, idx, text=True)
    assert open(os.path.join(d, "g1.config")).read().split() == ["200", "1500"]
    nv, ne, p, i = H.load_graph(d, "g1")
    assert (nv, ne) == (200, 1500) and np.array_equal(p, ptr) and np.array_equal(i, idx)
    # first text read leaves the binary caches behind, in the reference's layout (raw int32)
    assert np.array_equal(np.fromfile(os.path.join(d, "g1.graph.ptrdump"), np.int32), ptr)
    assert np.array_equal(np.fromfile(os.path.join(d, "g1.graph.edgedump"), np.int32), idx)
    os.remove(os.path.join(d, "g1.graph"))        # dumps alone are enough from now on
    nv, ne, p, i = H.load_graph(d, "g1")
    assert np.array_equal(p, ptr) and np.array_equal(i, idx)
    H.write_graph(d, "g2", ptr, idx, text=False)
    nv, ne, p, i = H.load_graph(d, "g2")
    assert np.array_equal(p, ptr) and np.array_equal(i, idx)
    with pytest.raises(H.SpmmB200Error) as e:
        H.load_graph(d, "missing")
    assert e.value.code == -3
    open(os.path.join(d, "bad.config"), "w").write("200 1499\n")     # ptr[num_v] != num_e (data.cu:40-45)
    np.asarray(ptr, np.int32).tofile(os.path.join(d, "bad.graph.ptrdump"))
    np.asarray(idx, np.int32).tofile(os.path.join(d, "bad.graph.edgedump"))
    with pytest.raises(H.SpmmB200Error):
        H.load_graph(d, "bad")


@pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built")
def test_loader_matches_reference_load_graph(tmp_path):
    """The reference's own load_graph (unmodified data.cu in oracle/_ref) reads what we write, and
    reads the same arrays we read. Host-only code, but the shim library links the CUDA runtime."""
    try:
        lib = refshim.lib()
    except OSError as e:      # CUDA runtime libraries not loadable on this machine
        pytest.skip(str(e))
    import ctypes as C
    ptr, idx = H.gen_graph(300, 2500, 120, seed=4)
    d = str(tmp_path)
    H.write_graph(d, "t", ptr, idx, text=True)
    nv, ne = C.c_int(0), C.c_int(0)
    p = np.empty(301, np.int32)
    i = np.empty(2500, np.int32)
    rc = lib.ref_load_graph(d.encode(), b"t", C.byref(nv), C.byref(ne), p.ctypes.data, i.ctypes.data, 301, 2500)
    assert rc == 0 and (nv.value, ne.value) == (300, 2500)
    assert np.array_equal(p, ptr) and np.array_equal(i, idx)
    # the dumps the reference wrote are the ones our loader reads
    os.remove(os.path.join(d, "t.graph"))
    _, _, p2, i2 = H.load_graph(d, "t")
    assert np.array_equal(p2, ptr) and np.array_equal(i2, idx)


def test_partition_rows():
    for name, parts in [("c0", 2), ("c0", 8), ("arxiv", 4), ("arxiv", 8)]:
        ptr, _ = H.gen_named_graph(name)
        b = H.partition_rows(ptr, parts)
        assert np.array_equal(b, P.partition_rows(ptr, parts))
        assert b[0] == 0 and b[-1] == len(ptr) - 1 and np.all(np.diff(b) >= 0)
        nnz = ptr[b[1:]].astype(np.int64) - ptr[b[:-1]]
        assert nnz.sum() == ptr[-1]
        # balanced to within the heaviest row
        assert nnz.max() - ptr[-1] / parts <= np.diff(ptr).max()
    # cost-balanced rule: nonzeros + row_cost per row (the plan's per-row overhead), same integer search
    for name, parts, rc in [("c0", 3, 1), ("arxiv", 8, 5), ("arxiv", 4, 64)]:
        ptr, _ = H.gen_named_graph(name)
        b = H.partition_rows(ptr, parts, rc)
        assert np.array_equal(b, P.partition_rows(ptr, parts, rc))
        cost = (ptr[b[1:]].astype(np.int64) - ptr[b[:-1]]) + rc * np.diff(b).astype(np.int64)
        assert cost.max() - (int(ptr[-1]) + rc * (len(ptr) - 1)) / parts <= np.diff(ptr).max() + rc
    assert H.plan_row_cost(232965, 114615892, 256) == 5 and H.plan_row_cost(2449029, 123718280, 256) == 1
    assert H.plan_row_cost(169343, 1166243, 32) == 1
    # degenerate: more parts than rows, empty graph
    ptr = np.asarray([0, 5, 5, 9], np.int32)
    assert np.array_equal(H.partition_rows(ptr, 8), P.partition_rows(ptr, 8))
    assert np.array_equal(H.partition_rows(np.zeros(1, np.int32), 3), [0, 0, 0, 0])
    assert np.array_equal(H.rebase_ptr(ptr, 1, 3), [0, 0, 4])


def test_plan_host_matches_oracle():
    for name, sl, ro in [("c0", 0, True), ("c0", 16, True), ("c0", 16, False), ("arxiv", 0, True), ("arxiv", 100, True)]:
        ptr, idx = H.gen_named_graph(name)
        val = np.zeros(len(idx), np.float32)
        got = H.plan_host(ptr, 256, sl, ro)
        want = P.plan(ptr, idx, val, sl or P.auto_seg_len(len(idx), 256), ro, pad=4)      # K=256: one lane group
        for k in ("row_perm", "heavy_rows", "heavy_seg0", "seg_desc"):
            assert np.array_equal(got[k], want[k]), (name, k)
        assert got["panel_len"] == len(want["panel"])
    # automatic cut-off: 256 when a full warp covers the feature slice (K >= 128), else 128..1024 by nnz
    assert P.auto_seg_len(114615892, 32) == 1024 and P.auto_seg_len(114615892, 64) == 1024 and P.auto_seg_len(20 << 20, 32) == 512
    ptr, idx = H.gen_named_graph("arxiv")
    assert P.auto_seg_len(len(idx), 256) == 256 and P.auto_seg_len(len(idx), 32) == 128 and P.auto_seg_len(1, 128) == 256
    assert P.auto_seg_len(len(idx), 100) == 256 and P.auto_seg_len(len(idx), 64) == 128
    a, b = H.plan_host(ptr, 32), P.plan(ptr, idx, np.zeros(len(idx), np.float32), 128, True, pad=16)   # K=32: 4 groups
    assert np.array_equal(a["seg_desc"], b["seg_desc"]) and np.array_equal(a["row_perm"], b["row_perm"])
    a, b = H.plan_host(ptr, 256), P.plan(ptr, idx, np.zeros(len(idx), np.float32), 256, True, pad=4)
    assert np.array_equal(a["seg_desc"], b["seg_desc"]) and np.array_equal(a["row_perm"], b["row_perm"])
    assert len(H.plan_host(ptr, 30)["heavy_rows"]) == 0       # scalar path keeps every row whole
    # cross-check with the student's task split (spmm_opt.cu:43-54): same number of pieces per row at 256
    from oracle import cpu as O
    ptr, _ = H.gen_named_graph("arxiv")
    assert P.student_split_check(ptr, O.student_tasks(ptr, 256))
    got = H.plan_host(ptr, 256, 256, True)
    deg = np.diff(ptr)
    per_row = np.bincount(got["seg_desc"][:, 0], minlength=len(deg))
    heavy = deg > 256
    assert np.array_equal(per_row[heavy], -(-deg[heavy] // 256))
    # every nonzero of a heavy row is covered exactly once, in order
    sd = got["seg_desc"]
    for r in got["heavy_rows"][:20]:
        s = sd[sd[:, 0] == r]
        assert s[0, 3] == ptr[r] and s[-1, 3] + s[-1, 2] == ptr[r + 1]
        assert np.all(s[1:, 3] == s[:-1, 3] + s[:-1, 2])
        assert s[:, 2].max() - s[:, 2].min() <= 1        # nnz-balanced


def test_pack_light_host_matches_oracle():
    rng = np.random.default_rng(3)
    for groups, steps in [(1, 256), (4, 64), (8, 32), (2, 7), (32, 32), (1, 1)]:
        for n in (0, 1, 5, 1000):
            cost = rng.integers(1, 3 * steps + 2, n).astype(np.int32)
            dst, ltask, length = H.pack_light_host(cost, groups, steps)
            odst, otask, olen = P.pack_light(cost, groups, steps)
            assert np.array_equal(dst, odst) and np.array_equal(ltask, otask) and length == olen, (groups, steps, n)
            if n:
                # slots of different rows never collide, tasks tile the panel, offsets stay 16-byte aligned
                used = np.concatenate([d + np.arange(c) * groups for d, c in zip(dst, cost)])
                assert len(np.unique(used)) == len(used) and used.max() < length
                assert np.all(ltask[:, 0] % 4 == 0) and np.all(ltask[:, 1] % 4 == 0)
                assert np.array_equal(ltask[1:, 0], (ltask[:, 0] + ltask[:, 1] * groups)[:-1])


def test_property_based_plan_and_packing():
    """hypothesis: random degree sequences / costs -> product == oracle, plus structural invariants."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(0, 700), min_size=0, max_size=300), st.integers(1, 300), st.booleans())
    def plan_case(degs, seg_len, reorder):
        ptr = np.concatenate([[0], np.cumsum(np.asarray(degs, np.int64))]).astype(np.int32)
        nnz = int(ptr[-1])
        idx = np.zeros(nnz, np.int32)
        val = np.zeros(nnz, np.float32)
        got = H.plan_host(ptr, 32, seg_len, reorder)
        want = P.plan(ptr, idx, val, seg_len, reorder, pad=16)       # feat 32 -> 8 lanes -> 4 groups -> pad 16
        for k in ("row_perm", "heavy_rows", "heavy_seg0", "seg_desc"):
            assert np.array_equal(got[k], want[k]), k
        # every row appears exactly once; segments tile their rows
        rows = np.concatenate([got["row_perm"], got["heavy_rows"]])
        assert sorted(rows.tolist()) == list(range(len(degs)))
        assert int(got["seg_desc"][:, 2].sum()) == int(sum(d for d in degs if d > seg_len))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 400), min_size=0, max_size=400), st.sampled_from([1, 2, 4, 8, 16, 32]), st.integers(1, 300))
    def pack_case(costs, groups, steps):
        cost = np.asarray(costs, np.int32)
        dst, ltask, length = H.pack_light_host(cost, groups, steps)
        odst, otask, olen = P.pack_light(cost, groups, steps)
        assert np.array_equal(dst, odst) and np.array_equal(ltask, otask) and length == olen
        if len(cost):
            used = np.concatenate([d + np.arange(c) * groups for d, c in zip(dst, cost)])
            assert len(np.unique(used)) == len(used) and used.max() < length and length % 2 == 0
            assert int((ltask[:, 1] * groups).sum()) == length

    plan_case()
    pack_case()
