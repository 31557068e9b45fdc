"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the parity oracle.

Two libraries live here:

* ``libem2oracle.so``  -- plain-C restatement of the reference hot path (``em2_oracle.c``), each
  function citing the reference file:line it follows.  Always buildable (``make -C oracle``).
* ``_ref/libem2ref.so`` -- the reference's OWN sources compiled unmodified from
  ``/root/reference/src`` plus ``ref_driver.cpp`` (``make -C oracle ref``).  Present whenever it was
  built in the authoring container; it travels to the GPU box as a prebuilt file.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker.  Nothing under ``expressionmatrix2_b200/``
imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libem2oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libem2ref.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    args = ["make", "-C", _HERE, "all"]
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ----------------------------------------------------------------------------------------------
# C restatement
# ----------------------------------------------------------------------------------------------
_olib = None


def olib():
    global _olib
    if _olib is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        L = C.CDLL(_ORACLE_SO)
        L.em2o_normal_stream.argtypes = [C.c_uint32, C.c_uint64, f64p]
        L.em2o_generate_lsh_vectors.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, f64p]
        L.em2o_cell_sums.argtypes = [C.c_uint64, u64p, f32p, f64p, f64p]
        L.em2o_signatures.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, f32p, f64p, f64p, C.c_uint64, u64p,
                                      C.c_double, u64p, f64p]
        L.em2o_similarity_table.argtypes = [C.c_uint64, f64p]
        L.em2o_mismatch_max.argtypes = [C.c_uint64, C.c_double]
        L.em2o_mismatch_max.restype = C.c_int64
        L.em2o_mismatch_counts.argtypes = [u64p, C.c_uint64, C.c_uint64, u32p, u32p, u32p]
        L.em2o_mismatch_row.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint32, u32p]
        L.em2o_mismatch_checksum.argtypes = [u64p, C.c_uint64, C.c_uint64, u64p, u64p]
        L.em2o_topk.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64,
                                u32p, f32p, u32p]
        L.em2o_topk.restype = C.c_double
        L.em2o_pair_loop.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, u64p, u64p]
        L.em2o_pair_loop.restype = C.c_double
        L.em2o_exact_similarities.argtypes = [C.c_uint64, u64p, u32p, f32p, f64p, f64p, C.c_uint64, u32p, u32p, f64p]
        L.em2o_exact_rows.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, f32p, f64p, f64p, C.c_uint64,
                                      C.c_uint64, f64p]
        L.em2o_murmur64a.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.em2o_murmur64a.restype = C.c_uint64
        _olib = L
    return _olib


def word_count(lsh_count: int) -> int:
    return (lsh_count - 1) // 64 + 1


def normal_stream(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float64)
    olib().em2o_normal_stream(seed, n, _ptr(out, f64p))
    return out


def generate_lsh_vectors(gene_count: int, lsh_count: int, seed: int) -> np.ndarray:
    U = np.empty((gene_count, lsh_count), np.float64)
    olib().em2o_generate_lsh_vectors(gene_count, lsh_count, seed, _ptr(U, f64p))
    return U


def cell_sums(toc, counts):
    toc = _c(toc, np.uint64)
    counts = _c(counts, np.float32)
    n = len(toc) - 1
    s1 = np.empty(n, np.float64)
    s2 = np.empty(n, np.float64)
    olib().em2o_cell_sums(n, _ptr(toc, u64p), _ptr(counts, f32p), _ptr(s1, f64p), _ptr(s2, f64p))
    return s1, s2


def signatures(toc, gene_ids, counts, sum1, U, eps: float = 1e-12, want_scalars: bool = False):
    toc = _c(toc, np.uint64)
    gene_ids = _c(gene_ids, np.uint32)
    counts = _c(counts, np.float32)
    sum1 = _c(sum1, np.float64)
    U = _c(U, np.float64)
    n = len(toc) - 1
    G, L = U.shape
    W = word_count(L)
    sig = np.zeros((n, W), np.uint64)
    nz = C.c_uint64(0)
    scal = np.empty((n, L), np.float64) if want_scalars else None
    olib().em2o_signatures(n, G, _ptr(toc, u64p), _ptr(gene_ids, u32p), _ptr(counts, f32p), _ptr(sum1, f64p),
                           _ptr(U, f64p), L, _ptr(sig, u64p), eps, C.byref(nz),
                           _ptr(scal, f64p) if want_scalars else None)
    if want_scalars:
        return sig, int(nz.value), scal
    return sig, int(nz.value)


def similarity_table(lsh_count: int) -> np.ndarray:
    t = np.empty(lsh_count + 1, np.float64)
    olib().em2o_similarity_table(lsh_count, _ptr(t, f64p))
    return t


def mismatch_max(lsh_count: int, threshold: float) -> int:
    return int(olib().em2o_mismatch_max(lsh_count, threshold))


def mismatch_counts(sig, c0, c1) -> np.ndarray:
    sig = _c(sig, np.uint64)
    c0 = _c(c0, np.uint32)
    c1 = _c(c1, np.uint32)
    out = np.empty(len(c0), np.uint32)
    olib().em2o_mismatch_counts(_ptr(sig, u64p), sig.shape[1], len(c0), _ptr(c0, u32p), _ptr(c1, u32p),
                                _ptr(out, u32p))
    return out


def mismatch_row(sig, cell0: int) -> np.ndarray:
    sig = _c(sig, np.uint64)
    out = np.empty(sig.shape[0], np.uint32)
    olib().em2o_mismatch_row(_ptr(sig, u64p), sig.shape[1], sig.shape[0], cell0, _ptr(out, u32p))
    return out


def mismatch_checksum(sig):
    sig = _c(sig, np.uint64)
    a = C.c_uint64(0)
    b = C.c_uint64(0)
    olib().em2o_mismatch_checksum(_ptr(sig, u64p), sig.shape[1], sig.shape[0], C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def topk(sig, lsh_count: int, k: int, threshold: float, row_begin: int = 0, row_end: int | None = None):
    """Deterministic top-k oracle. Returns (ids [R,k] u32, sims [R,k] f32, used [R] u32, seconds)."""
    sig = _c(sig, np.uint64)
    n = sig.shape[0]
    row_end = n if row_end is None else row_end
    R = row_end - row_begin
    ids = np.zeros((R, k), np.uint32)
    sims = np.zeros((R, k), np.float32)
    used = np.zeros(R, np.uint32)
    t = olib().em2o_topk(_ptr(sig, u64p), n, lsh_count, k, threshold, row_begin, row_end, _ptr(ids, u32p),
                         _ptr(sims, f32p), _ptr(used, u32p))
    return ids, sims, used, float(t)


def pair_loop(sig, lsh_count: int, threshold: float, row_begin: int, row_end: int):
    sig = _c(sig, np.uint64)
    pairs = C.c_uint64(0)
    passed = C.c_uint64(0)
    t = olib().em2o_pair_loop(_ptr(sig, u64p), sig.shape[0], lsh_count, threshold, row_begin, row_end,
                              C.byref(pairs), C.byref(passed))
    return float(t), int(pairs.value), int(passed.value)


def exact_similarities(gene_count, toc, gene_ids, counts, sum1, sum2, c0, c1) -> np.ndarray:
    toc = _c(toc, np.uint64)
    gene_ids = _c(gene_ids, np.uint32)
    counts = _c(counts, np.float32)
    sum1 = _c(sum1, np.float64)
    sum2 = _c(sum2, np.float64)
    c0 = _c(c0, np.uint32)
    c1 = _c(c1, np.uint32)
    out = np.empty(len(c0), np.float64)
    olib().em2o_exact_similarities(gene_count, _ptr(toc, u64p), _ptr(gene_ids, u32p), _ptr(counts, f32p),
                                   _ptr(sum1, f64p), _ptr(sum2, f64p), len(c0), _ptr(c0, u32p), _ptr(c1, u32p),
                                   _ptr(out, f64p))
    return out


def exact_rows(gene_count, toc, gene_ids, counts, sum1, sum2, row_begin, row_end) -> np.ndarray:
    toc = _c(toc, np.uint64)
    gene_ids = _c(gene_ids, np.uint32)
    counts = _c(counts, np.float32)
    sum1 = _c(sum1, np.float64)
    sum2 = _c(sum2, np.float64)
    n = len(toc) - 1
    out = np.empty((row_end - row_begin, n), np.float64)
    olib().em2o_exact_rows(n, gene_count, _ptr(toc, u64p), _ptr(gene_ids, u32p), _ptr(counts, f32p),
                           _ptr(sum1, f64p), _ptr(sum2, f64p), row_begin, row_end, _ptr(out, f64p))
    return out


def exact_topk(gene_count, toc, gene_ids, counts, k: int, threshold: float, row_begin: int = 0,
               row_end: int | None = None):
    """Deterministic top-k of the exact path (findSimilarPairs0, src/ExpressionMatrixFindSimilarPairs.cpp:60-80 +
    SimilarPairs::add/sort): per cell, among other cells with r > threshold (double compare), the k largest
    by (float(r) desc, cell id asc) -- SimilarPairs stores float similarities (src/SimilarPairs.hpp:53-56) and
    sort() orders by (similarity desc, id asc) (src/orderPairs.hpp:44-52).
    Returns (ids uint32[rows,k], sims float32[rows,k], used uint32[rows], r float64[rows,N])."""
    n = len(toc) - 1
    row_end = n if row_end is None else row_end
    s1, s2 = cell_sums(toc, counts)
    r = exact_rows(gene_count, toc, gene_ids, counts, s1, s2, row_begin, row_end)
    rows = row_end - row_begin
    ids = np.zeros((rows, k), np.uint32)
    sims = np.zeros((rows, k), np.float32)
    used = np.zeros(rows, np.uint32)
    cols = np.arange(n)
    for i in range(rows):
        row = r[i]
        with np.errstate(invalid="ignore"):
            ok = (row > threshold) & (cols != row_begin + i)
        cand = cols[ok]
        f = row[ok].astype(np.float32)
        order = np.lexsort((cand, -f.astype(np.float64)))[:k]
        u = len(order)
        ids[i, :u] = cand[order]
        sims[i, :u] = f[order]
        used[i] = u
    return ids, sims, used, r


def subset(toc, gene_ids, counts, gene_count: int, gene_set, cell_set):
    """ExpressionMatrixSubset construction (reference src/ExpressionMatrixSubset.cpp:9-42), restated: for every
    cell of the sorted `cell_set`, the stored counts whose gene is in the sorted `gene_set`, in stored order,
    with gene ids replaced by their index in `gene_set` (GeneSet::getLocalGeneId, src/GeneSet.hpp:70-77).
    Returns (toc uint64[n+1], local_gene_ids uint32[nnz'], counts float32[nnz'])."""
    toc = _c(toc, np.uint64)
    gene_ids = _c(gene_ids, np.uint32)
    counts = _c(counts, np.float32)
    gene_set = _c(gene_set, np.uint32)
    cell_set = _c(cell_set, np.uint32)
    local = np.full(gene_count, 0xFFFFFFFF, np.uint32)
    local[gene_set] = np.arange(len(gene_set), dtype=np.uint32)
    out_toc = np.zeros(len(cell_set) + 1, np.uint64)
    gs, cs = [], []
    n = 0
    for i, c in enumerate(cell_set):
        b, e = int(toc[c]), int(toc[c + 1])
        l = local[gene_ids[b:e]]
        keep = l != 0xFFFFFFFF
        gs.append(l[keep])
        cs.append(counts[b:e][keep])
        n += int(keep.sum())
        out_toc[i + 1] = n
    g = np.concatenate(gs) if gs else np.zeros(0, np.uint32)
    c = np.concatenate(cs) if cs else np.zeros(0, np.float32)
    return out_toc, g.astype(np.uint32), c.astype(np.float32)


def ref_subset(toc, gene_ids, counts, gene_count: int, gene_set, cell_set):
    """The reference's own ExpressionMatrixSubset constructor (oracle/_ref).  Returns (toc, genes, counts, sum1, sum2)."""
    import tempfile
    toc = _c(toc, np.uint64)
    gene_ids = _c(gene_ids, np.uint32)
    counts = _c(counts, np.float32)
    gene_set = _c(gene_set, np.uint32)
    cell_set = _c(cell_set, np.uint32)
    n = len(cell_set)
    out_toc = np.zeros(n + 1, np.uint64)
    out_g = np.zeros(max(1, len(gene_ids)), np.uint32)
    out_c = np.zeros(max(1, len(gene_ids)), np.float32)
    s1 = np.zeros(n, np.float64)
    s2 = np.zeros(n, np.float64)
    nnz = C.c_uint64(0)
    L = rlib()
    L.em2ref_subset.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, u64p, u32p, f32p, C.c_uint64, u32p, C.c_uint64, u32p,
                                u64p, u32p, f32p, C.POINTER(C.c_uint64), f64p, f64p]
    with tempfile.TemporaryDirectory(prefix="em2ref-") as d:
        _check(L.em2ref_subset(d.encode(), len(toc) - 1, gene_count, _ptr(toc, u64p), _ptr(gene_ids, u32p),
                               _ptr(counts, f32p), len(gene_set), _ptr(gene_set, u32p), n, _ptr(cell_set, u32p),
                               _ptr(out_toc, u64p), _ptr(out_g, u32p), _ptr(out_c, f32p), C.byref(nnz),
                               _ptr(s1, f64p), _ptr(s2, f64p)))
    m = int(nnz.value)
    return out_toc, out_g[:m].copy(), out_c[:m].copy(), s1, s2


def cell_graph_edges(ids, sims, used, vertex_of, similarity_threshold: float, max_connectivity: int):
    """The edge loop of CellGraph::CellGraph (reference src/CellGraph.cpp:60-107), restated literally over a
    SimilarPairs payload: cells in order; per cell walk the row, `break` at the first similarity < threshold (the
    stored float promoted to double against the double parameter, CellGraph.cpp:84), skip neighbours that are not
    vertices, stop when the kept neighbours number max_connectivity (tested after the push_back, so 0 never stops);
    add the edge unless it exists.  vertex_of[c] = vertex index or 0xFFFFFFFF.
    Returns (vertex0 uint32[E], vertex1 uint32[E], similarity float32[E]) in insertion order.
    (The reference's CellGraph itself needs Boost.Graph and cannot be compiled here: restatement only.)"""
    ids = np.asarray(ids)
    sims = np.asarray(sims, np.float32)
    thr = float(similarity_threshold)
    seen = set()
    v0s, v1s, ss = [], [], []
    for c in range(len(used)):
        v0 = int(vertex_of[c])
        if v0 == 0xFFFFFFFF:
            continue
        kept = []
        for i in range(int(used[c])):
            if float(sims[c, i]) < thr:
                break
            v1 = int(vertex_of[int(ids[c, i])])
            if v1 == 0xFFFFFFFF:
                continue
            kept.append((v1, sims[c, i]))
            if len(kept) == max_connectivity:
                break
        for v1, sim in kept:
            key = (min(v0, v1), max(v0, v1))
            if key in seen:
                continue
            seen.add(key)
            v0s.append(v0)
            v1s.append(v1)
            ss.append(sim)
    return np.array(v0s, np.uint32), np.array(v1s, np.uint32), np.array(ss, np.float32)


def signature_graph(signatures, lsh_count: int, min_cell_count: int):
    """Vertices and edges of the reference's SignatureGraph, restated in numpy/Python:
    ExpressionMatrix::createSignatureGraph (src/ExpressionMatrixSignatureGraph.cpp:69-75, 111-125): cells with the same
    signature -> one vertex, std::map order = lexicographic over the 64-bit words (src/BitSet.hpp:157-160), cells in
    ascending id, signatures with fewer than min_cell_count cells dropped; SignatureGraph::createEdges
    (src/SignatureGraph.cpp:23-48): per vertex, per ZERO bit in bit order (bit 0 = MSB of word 0, BitSet.hpp:57-63),
    an edge to the vertex whose signature has that bit set.
    Returns (cell_order uint32, vertex_offsets uint64[V+1], edges int64[E, 2])."""
    sig = np.ascontiguousarray(signatures, np.uint64)
    n, W = sig.shape
    groups = {}
    for c in range(n):
        groups.setdefault(tuple(int(x) for x in sig[c]), []).append(c)
    keys = sorted(k for k, cells in groups.items() if len(cells) >= min_cell_count)      # tuple order == word order
    index = {k: v for v, k in enumerate(keys)}
    order, offsets = [], [0]
    for k in keys:
        order.extend(groups[k])
        offsets.append(len(order))
    edges = []
    for v0, k in enumerate(keys):
        for bit in range(lsh_count):
            w, mask = bit >> 6, 1 << (63 - (bit & 63))
            if k[w] & mask:
                continue
            k1 = k[:w] + (k[w] | mask,) + k[w + 1:]
            v1 = index.get(k1)
            if v1 is not None:
                edges.append((v0, v1))
    return (np.array(order, np.uint32), np.array(offsets, np.uint64),
            np.array(edges, np.int64).reshape(-1, 2))


def ref_signature_graph(signatures, lsh_count: int, min_cell_count: int):
    """The same through the reference's own BitSet classes (oracle/_ref, ref_driver.cpp em2ref_signature_graph)."""
    sig = np.ascontiguousarray(signatures, np.uint64)
    n = sig.shape[0]
    L = rlib()
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.em2ref_signature_graph.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, u32p, u64p, u64p, u32p, u32p, C.c_uint64, u64p]
    L.em2ref_signature_graph.restype = C.c_int
    order = np.zeros(max(n, 1), np.uint32)
    offsets = np.zeros(n + 1, np.uint64)
    cap = max(1, n * lsh_count)
    e0, e1 = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    vc, ec = C.c_uint64(0), C.c_uint64(0)
    _check(L.em2ref_signature_graph(_ptr(sig, u64p), n, lsh_count, min_cell_count, _ptr(order, u32p), _ptr(offsets, u64p),
                                    C.byref(vc), _ptr(e0, u32p), _ptr(e1, u32p), cap, C.byref(ec)))
    v, e = int(vc.value), int(ec.value)
    return order[: int(offsets[v])].copy(), offsets[: v + 1].copy(), np.stack([e0[:e], e1[:e]], axis=1).astype(np.int64)


def murmur64a(data: bytes, seed: int = 231) -> int:
    buf = C.create_string_buffer(data, len(data))
    return int(olib().em2o_murmur64a(buf, len(data), seed))


# ----------------------------------------------------------------------------------------------
# The reference itself (oracle/_ref)
# ----------------------------------------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(_REF_SO)


_rlib = None


def rlib():
    global _rlib
    if _rlib is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libem2ref.so is not built (run `make -C oracle ref` where "
                               "/root/reference exists)")
        L = C.CDLL(_REF_SO)
        L.em2ref_last_error.restype = C.c_char_p
        L.em2ref_open_csr.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, u64p, u32p, f32p, C.c_uint64, C.c_uint32,
                                      C.POINTER(C.c_void_p)]
        L.em2ref_open_signatures.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, u64p, C.POINTER(C.c_void_p)]
        L.em2ref_close.argtypes = [C.c_void_p]
        L.em2ref_word_count.argtypes = [C.c_void_p]
        L.em2ref_word_count.restype = C.c_uint64
        L.em2ref_signature_seconds.argtypes = [C.c_void_p]
        L.em2ref_signature_seconds.restype = C.c_double
        L.em2ref_lsh_ctor_seconds.argtypes = [C.c_void_p]
        L.em2ref_lsh_ctor_seconds.restype = C.c_double
        L.em2ref_get_signatures.argtypes = [C.c_void_p, u64p]
        L.em2ref_get_lsh_vectors.argtypes = [C.c_void_p, f64p]
        L.em2ref_get_sums.argtypes = [C.c_void_p, f64p, f64p]
        L.em2ref_get_similarity_table.argtypes = [C.c_void_p, f64p]
        L.em2ref_mismatch_counts.argtypes = [C.c_void_p, C.c_uint64, u32p, u32p, u32p]
        L.em2ref_mismatch_row.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.em2ref_exact_similarity.argtypes = [C.c_void_p, C.c_uint64, u32p, u32p, f64p]
        L.em2ref_find_similar_pairs4_loop.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint32, C.c_uint32,
                                                      u32p, f32p, u32p, f64p, u64p]
        L.em2ref_topk_deterministic.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint32, C.c_uint32,
                                                u32p, f32p, u32p, f64p]
        L.em2ref_write_similar_pairs.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                                 u32p, f32p, u32p]
        L.em2ref_read_similar_pairs.argtypes = [C.c_char_p, C.c_char_p, u64p, u64p, u32p, f32p, u32p]
        L.em2ref_keep_best_less.argtypes = [C.c_uint64, i64p, C.c_uint64, i64p, u64p]
        L.em2ref_normal_stream.argtypes = [C.c_uint32, C.c_uint64, f64p]
        L.em2ref_murmur64a.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.em2ref_murmur64a.restype = C.c_uint64
        _rlib = L
    return _rlib


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("reference oracle: " + rlib().em2ref_last_error().decode())


class Reference:
    """The reference's own Lsh / ExpressionMatrixSubset / SimilarPairs objects behind a handle."""

    def __init__(self, handle, tmpdir, cell_count, lsh_count, gene_count=None):
        self._h = handle
        self._tmp = tmpdir
        self.cell_count = cell_count
        self.lsh_count = lsh_count
        self.gene_count = gene_count

    @classmethod
    def from_csr(cls, toc, gene_ids, counts, gene_count: int, lsh_count: int, seed: int = 231) -> "Reference":
        toc = _c(toc, np.uint64)
        gene_ids = _c(gene_ids, np.uint32)
        counts = _c(counts, np.float32)
        tmp = tempfile.TemporaryDirectory(prefix="em2ref-")
        h = C.c_void_p()
        _check(rlib().em2ref_open_csr(tmp.name.encode(), len(toc) - 1, gene_count, _ptr(toc, u64p),
                                      _ptr(gene_ids, u32p), _ptr(counts, f32p), lsh_count, seed, C.byref(h)))
        return cls(h, tmp, len(toc) - 1, lsh_count, gene_count)

    @classmethod
    def from_signatures(cls, sig, lsh_count: int) -> "Reference":
        sig = _c(sig, np.uint64)
        tmp = tempfile.TemporaryDirectory(prefix="em2ref-")
        h = C.c_void_p()
        _check(rlib().em2ref_open_signatures(tmp.name.encode(), sig.shape[0], lsh_count, _ptr(sig, u64p),
                                             C.byref(h)))
        return cls(h, tmp, sig.shape[0], lsh_count)

    def close(self):
        if self._h is not None:
            _check(rlib().em2ref_close(self._h))
            self._h = None
            self._tmp.cleanup()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def signature_seconds(self) -> float:
        return float(rlib().em2ref_signature_seconds(self._h))

    @property
    def lsh_ctor_seconds(self) -> float:
        return float(rlib().em2ref_lsh_ctor_seconds(self._h))

    def signatures(self) -> np.ndarray:
        W = int(rlib().em2ref_word_count(self._h))
        out = np.empty((self.cell_count, W), np.uint64)
        _check(rlib().em2ref_get_signatures(self._h, _ptr(out, u64p)))
        return out

    def lsh_vectors(self) -> np.ndarray:
        out = np.empty((self.gene_count, self.lsh_count), np.float64)
        _check(rlib().em2ref_get_lsh_vectors(self._h, _ptr(out, f64p)))
        return out

    def sums(self):
        s1 = np.empty(self.cell_count, np.float64)
        s2 = np.empty(self.cell_count, np.float64)
        _check(rlib().em2ref_get_sums(self._h, _ptr(s1, f64p), _ptr(s2, f64p)))
        return s1, s2

    def similarity_table(self) -> np.ndarray:
        out = np.empty(self.lsh_count + 1, np.float64)
        _check(rlib().em2ref_get_similarity_table(self._h, _ptr(out, f64p)))
        return out

    def find_similar_pairs7(self, k: int, similarity_threshold: float, slice_lengths, max_check: int, log2_bucket_count: int):
        """ExpressionMatrix::findSimilarPairs7 (bucketed LSH; src/ExpressionMatrixLsh.cpp:507-827) restated over this
        reference Lsh object (ref_driver.cpp em2ref_find_similar_pairs7).  Returns (ids, sims float32, used)."""
        L = rlib()
        i32p = C.POINTER(C.c_int32)
        L.em2ref_find_similar_pairs7.argtypes = [C.c_void_p, C.c_uint64, C.c_double, i32p, C.c_uint64, C.c_uint32, C.c_uint64,
                                                 u32p, f32p, u32p]
        L.em2ref_find_similar_pairs7.restype = C.c_int
        sl = np.ascontiguousarray(slice_lengths, np.int32)
        ids = np.zeros((self.cell_count, k), np.uint32)
        sims = np.zeros((self.cell_count, k), np.float32)
        used = np.zeros(self.cell_count, np.uint32)
        _check(L.em2ref_find_similar_pairs7(self._h, k, similarity_threshold, _ptr(sl, i32p), len(sl), max_check, log2_bucket_count,
                                            _ptr(ids, u32p), _ptr(sims, f32p), _ptr(used, u32p)))
        return ids, sims, used

    def mismatch_counts(self, c0, c1) -> np.ndarray:
        c0 = _c(c0, np.uint32)
        c1 = _c(c1, np.uint32)
        out = np.empty(len(c0), np.uint32)
        _check(rlib().em2ref_mismatch_counts(self._h, len(c0), _ptr(c0, u32p), _ptr(c1, u32p), _ptr(out, u32p)))
        return out

    def mismatch_row(self, cell0: int) -> np.ndarray:
        out = np.empty(self.cell_count, np.uint32)
        _check(rlib().em2ref_mismatch_row(self._h, cell0, _ptr(out, u32p)))
        return out

    def exact_similarity(self, c0, c1) -> np.ndarray:
        c0 = _c(c0, np.uint32)
        c1 = _c(c1, np.uint32)
        out = np.empty(len(c0), np.float64)
        _check(rlib().em2ref_exact_similarity(self._h, len(c0), _ptr(c0, u32p), _ptr(c1, u32p), _ptr(out, f64p)))
        return out

    def find_similar_pairs4_loop(self, k: int, threshold: float, row_begin: int = 0, row_end: int | None = None,
                                 want_pairs: bool = True):
        """Literal findSimilarPairs4 pair loop. Returns dict(ids, sims, used, seconds, pairs)."""
        n = self.cell_count
        row_end = n if row_end is None else row_end
        secs = C.c_double(0)
        visited = C.c_uint64(0)
        if want_pairs:
            ids = np.zeros((n, k), np.uint32)
            sims = np.zeros((n, k), np.float32)
            used = np.zeros(n, np.uint32)
            _check(rlib().em2ref_find_similar_pairs4_loop(self._h, k, threshold, row_begin, row_end, _ptr(ids, u32p),
                                                          _ptr(sims, f32p), _ptr(used, u32p), C.byref(secs),
                                                          C.byref(visited)))
        else:
            ids = sims = used = None
            _check(rlib().em2ref_find_similar_pairs4_loop(self._h, k, threshold, row_begin, row_end, None, None,
                                                          None, C.byref(secs), C.byref(visited)))
        return dict(ids=ids, sims=sims, used=used, seconds=float(secs.value), pairs=int(visited.value))

    def topk_deterministic(self, k: int, threshold: float, row_begin: int = 0, row_end: int | None = None):
        n = self.cell_count
        row_end = n if row_end is None else row_end
        R = row_end - row_begin
        ids = np.zeros((R, k), np.uint32)
        sims = np.zeros((R, k), np.float32)
        used = np.zeros(R, np.uint32)
        secs = C.c_double(0)
        _check(rlib().em2ref_topk_deterministic(self._h, k, threshold, row_begin, row_end, _ptr(ids, u32p),
                                                _ptr(sims, f32p), _ptr(used, u32p), C.byref(secs)))
        return ids, sims, used, float(secs.value)


def ref_normal_stream(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float64)
    _check(rlib().em2ref_normal_stream(seed, n, _ptr(out, f64p)))
    return out


def ref_murmur64a(data: bytes, seed: int = 231) -> int:
    buf = C.create_string_buffer(data, len(data))
    return int(rlib().em2ref_murmur64a(buf, len(data), seed))


def ref_keep_best_less(values, k: int) -> np.ndarray:
    v = _c(values, np.int64)
    out = np.empty(len(v), np.int64)
    cnt = C.c_uint64(0)
    _check(rlib().em2ref_keep_best_less(len(v), _ptr(v, i64p), k, _ptr(out, i64p), C.byref(cnt)))
    return out[: cnt.value]


def ref_write_similar_pairs(directory: str, name: str, gene_count: int, ids, sims, used) -> None:
    ids = _c(ids, np.uint32)
    sims = _c(sims, np.float32)
    used = _c(used, np.uint32)
    n, k = ids.shape
    _check(rlib().em2ref_write_similar_pairs(directory.encode(), name.encode(), n, gene_count, k, _ptr(ids, u32p),
                                             _ptr(sims, f32p), _ptr(used, u32p)))


def ref_cell_graph_edges(ids, sims, used, cell_set, similarity_threshold: float, max_connectivity: int, gene_count: int = 4):
    """Edges of the reference's OWN CellGraph constructor (src/CellGraph.cpp:33-117, compiled unmodified into
    oracle/_ref against the Boost.Graph stand-in of boost_shim/) for a SimilarPairs payload over cells 0..N-1 and the
    graph cell set `cell_set` (sorted cell ids).  Returns (vertex0, vertex1, similarity) in the graph's edge order,
    vertices as positions in cell_set."""
    cell_set = _c(cell_set, np.uint32)
    with tempfile.TemporaryDirectory(prefix="em2ref-") as d:
        ref_write_similar_pairs(d, "P", gene_count, ids, sims, used)
        cap = max(1, len(cell_set) * ids.shape[1])
        v0 = np.zeros(cap, np.uint32)
        v1 = np.zeros(cap, np.uint32)
        ss = np.zeros(cap, np.float32)
        n = C.c_uint64(0)
        rlib().em2ref_cell_graph_edges.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, u32p, C.c_double, C.c_uint64, C.c_uint64,
                                                   u32p, u32p, f32p, u64p]
        _check(rlib().em2ref_cell_graph_edges(d.encode(), b"P", len(cell_set), _ptr(cell_set, u32p), similarity_threshold,
                                              max_connectivity, cap, _ptr(v0, u32p), _ptr(v1, u32p), _ptr(ss, f32p), C.byref(n)))
    m = int(n.value)
    return v0[:m].copy(), v1[:m].copy(), ss[:m].copy()


def ref_read_similar_pairs(directory: str, name: str):
    k = C.c_uint64(0)
    n = C.c_uint64(0)
    _check(rlib().em2ref_read_similar_pairs(directory.encode(), name.encode(), C.byref(k), C.byref(n), None, None,
                                            None))
    ids = np.zeros((n.value, k.value), np.uint32)
    sims = np.zeros((n.value, k.value), np.float32)
    used = np.zeros(n.value, np.uint32)
    _check(rlib().em2ref_read_similar_pairs(directory.encode(), name.encode(), C.byref(k), C.byref(n),
                                            _ptr(ids, u32p), _ptr(sims, f32p), _ptr(used, u32p)))
    return ids, sims, used
