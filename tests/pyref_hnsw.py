"""SECOND, INDEPENDENT restatement of the reference's HNSW path — pure Python, small cases only (test infrastructure).

oracle/hnsw_oracle.cpp is the checker the GPU tests use; the reference itself (Rust) cannot be compiled in this image,
so the oracle is pinned on the reference's own known-answer tests.  This file adds a second pin: a line-by-line
transliteration of src/hnsw/core.rs written without looking at the oracle's code, including Rust's
std::collections::BinaryHeap (whose sift rules decide the order of equal sims).  tests/test_pyref_cross_check.py
requires the two restatements to agree on graphs (every list, in order), touched sets, deletes and search results —
on continuous data AND on grid data full of ties.

Every function cites the reference lines it follows.  Nodes are integer ids (the reference uses name strings; identity
is all that matters on this path).
"""
import math

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------- std::collections::BinaryHeap (max-heap on `le`)
class RustHeap:
    """Array-backed binary max-heap with the sift rules of Rust's std (library/alloc/src/collections/binary_heap.rs):
    push = append + sift_up (moves up while NOT elem <= parent); pop = swap the last element into slot 0, then
    sift_down_to_bottom (walk the hole to a leaf taking the right child iff left <= right) followed by sift_up.
    `reverse=True` models BinaryHeap<Reverse<T>> (std::cmp::Reverse swaps the comparison).  Items are (sim, node);
    SimPair compares by sim only (core.rs:271-300)."""

    def __init__(self, reverse=False, data=None):
        self.reverse = reverse
        self.data = list(data) if data is not None else []

    def _le(self, a, b):                       # a <= b in the heap's order
        return (b[0] <= a[0]) if self.reverse else (a[0] <= b[0])

    def __len__(self):
        return len(self.data)

    def clone(self):
        return RustHeap(self.reverse, self.data)

    def into_vec(self):                        # BinaryHeap::into_vec / IntoIterator: the raw array order
        return list(self.data)

    def peek(self):
        return self.data[0]

    def _sift_up(self, start, pos):
        d = self.data
        elem = d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if self._le(elem, d[parent]):
                break
            d[pos] = d[parent]
            pos = parent
        d[pos] = elem

    def push(self, item):
        self.data.append(item)
        self._sift_up(0, len(self.data) - 1)

    def _sift_down_to_bottom(self, pos):
        d = self.data
        end = len(d)
        start = pos
        elem = d[pos]
        child = 2 * pos + 1
        while child <= max(end - 2, 0) and child + 1 < end:
            if self._le(d[child], d[child + 1]):
                child += 1
            d[pos] = d[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1:
            d[pos] = d[child]
            pos = child
        d[pos] = elem
        self._sift_up(start, pos)

    def pop(self):
        item = self.data.pop()
        if self.data:
            item, self.data[0] = self.data[0], item
            self._sift_down_to_bottom(0)
        return item


# ---------------------------------------------------------------- metrics.rs:79-84 (the scalar path; dim % 32 != 0)
def sim_func_euc(a, b):
    """-(sum (a_i - b_i)^2), a strict left fold in f32 with separately rounded multiply and add (metrics.rs:79-84)."""
    acc = F32(0.0)
    for x, y in zip(a, b):
        d = F32(x) - F32(y)
        acc = F32(acc + F32(d * d))
    return F32(-acc)


# ---------------------------------------------------------------- metrics.rs:25-77 (the AVX2 + FMA path; dim % 32 == 0)
def _round_f32(fr):
    """Correctly rounded (nearest, ties to even) f32 of an exact rational; normal range only."""
    from fractions import Fraction

    if fr == 0:
        return F32(0.0)
    neg = fr < 0
    fr = -fr if neg else fr
    e = fr.numerator.bit_length() - fr.denominator.bit_length()
    if Fraction(2) ** e > fr:
        e -= 1
    assert Fraction(2) ** e <= fr < Fraction(2) ** (e + 1) and -126 <= e <= 127
    scaled = fr / Fraction(2) ** (e - 23)          # in [2^23, 2^24)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    v = float(n) * 2.0 ** (e - 23)                 # exact in f64
    return F32(-v if neg else v)


def _fma_f32(a, b, c):
    """_mm256_fmadd_ps lane: a * b + c with ONE rounding (exact rational arithmetic, then round to f32)."""
    from fractions import Fraction

    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def sim_func_avx_euc(a, b):
    """metrics.rs:48-77: four 8-lane accumulators fed by fused multiply-adds over blocks of 32 floats, then
    (euc1 + euc2) + (euc3 + euc4) (:71-74), low 128 + high 128 (:37-39), and hsum_ps_sse3 (:25-32): (v0 + v1) + (v2 + v3)."""
    n = len(a)
    assert n % 32 == 0
    euc = [[F32(0.0)] * 8 for _ in range(4)]
    for i in range(0, n, 32):
        for k in range(4):
            for j in range(8):
                d = F32(F32(a[i + 8 * k + j]) - F32(b[i + 8 * k + j]))
                euc[k][j] = _fma_f32(d, d, euc[k][j])
    v = [F32(F32(euc[0][j] + euc[1][j]) + F32(euc[2][j] + euc[3][j])) for j in range(8)]
    lo = [F32(v[j] + v[j + 4]) for j in range(4)]
    return F32(-F32(F32(lo[0] + lo[1]) + F32(lo[2] + lo[3])))


class _Node:
    def __init__(self, data):
        self.data = np.asarray(data, dtype=F32)
        self.neighbors = []                    # Vec<Vec<NodeWeak>>: neighbors[level] = ordered list of ids

    def push_levels(self, level):              # core.rs:127-135
        while len(self.neighbors) < level + 1:
            self.neighbors.append([])

    def add_neighbor(self, level, n):          # core.rs:137-143
        self.push_levels(level)
        if n not in self.neighbors[level]:
            self.neighbors[level].append(n)

    def rm_neighbor(self, level, n):           # core.rs:145-152 (panics if absent)
        self.neighbors[level].remove(n)


class PyRefIndex:
    """hnsw::Index<f32, f32> (core.rs:302-346) with injected levels instead of the entropy-seeded RNG (core.rs:344)."""

    def __init__(self, dim, m, ef_construction):
        assert dim % 32 != 0, "this restatement only carries the scalar metric path"
        self.data_dim, self.m, self.m_max, self.m_max_0 = dim, m, m, 2 * m      # core.rs:331-336
        self.ef_construction = ef_construction
        self.level_mult = 1.0 / math.log(1.0 * m)                                # core.rs:338
        self.node_count, self.max_layer = 0, 0
        self.layers, self.nodes, self.enterpoint = [], {}, None
        self.levels = {}                       # id -> the level the node was inserted with (for exports)
        self.next_id = 0
        self.last_updated = set()
        self.n_dist = 0

    def mfunc(self, a, b):
        self.n_dist += 1
        return sim_func_euc(a, b)

    # ---- core.rs:383-412
    def add_node(self, data, level):
        assert len(data) == self.data_dim
        nid = self.next_id
        self.next_id += 1
        self.last_updated = set()
        if self.node_count == 0:               # core.rs:393-405: enterpoint, layer 0, no level draw
            self.nodes[nid] = _Node(data)
            self.enterpoint = nid
            self.layers.append({nid})
            self.node_count += 1
            self.levels[nid] = 0
            return nid
        self.insert(nid, data, level)
        return nid

    # ---- core.rs:489-599
    def insert(self, nid, data, l):
        l_max = self.max_layer
        self.nodes[nid] = _Node(data)
        self.levels[nid] = l
        self.node_count += 1
        query = nid
        data = self.nodes[nid].data
        ep = self.enterpoint
        lc = l_max
        while lc > l:                          # core.rs:511-520
            w = self.search_level(data, ep, 1, lc)
            ep = w.pop()[1]
            if lc == 0:
                break
            lc -= 1
        updated = set()
        for lc in range(min(l_max, l), -1, -1):                                   # core.rs:523
            w = self.search_level(data, ep, self.ef_construction, lc)
            neighbors = self.select_neighbors(query, w, self.m, lc, None)          # core.rs:531
            self.connect_neighbors(query, neighbors, lc)
            for _, n in neighbors.into_vec():                                      # core.rs:535-537
                updated.add(n)
            while len(neighbors):                                                  # core.rs:540-574
                esim, e = neighbors.pop()
                econn = RustHeap()
                for n in self.nodes[e].neighbors[lc]:
                    econn.push((self.mfunc(self.nodes[e].data, self.nodes[n].data), n))
                m_max = self.m_max_0 if lc == 0 else self.m_max
                if len(econn) > m_max:
                    enewconn = self.select_neighbors(e, econn, m_max, lc, None)
                    updated |= self.update_node_connections(e, enewconn, econn, lc, None)
            ep = w.peek()[1]                                                       # core.rs:576
        self.last_updated = updated                                                # core.rs:580-584
        if l > l_max:                                                              # core.rs:587-593
            self.max_layer = l
            self.enterpoint = query
            while len(self.layers) < l + 1:
                self.layers.append(set())
        self.layers[l].add(query)                                                  # core.rs:596

    # ---- core.rs:607-675
    def search_level(self, query, ep, ef, level):
        v = {ep}
        qsim = self.mfunc(query, self.nodes[ep].data)
        c = RustHeap()
        w = RustHeap(reverse=True)
        c.push((qsim, ep))
        w.push((qsim, ep))
        while len(c):
            cpair = c.pop()
            fpair = w.peek()
            if cpair[0] < fpair[0]:                                                # core.rs:635
                break
            self.nodes[cpair[1]].push_levels(level)                                # core.rs:642
            for neighbor in self.nodes[cpair[1]].neighbors[level]:                 # list order, core.rs:646
                if neighbor not in v:
                    v.add(neighbor)
                    fpair = w.peek()                                               # core.rs:651
                    esim = self.mfunc(query, self.nodes[neighbor].data)
                    if esim > fpair[0] or len(w) < ef:                             # core.rs:657
                        c.push((esim, neighbor))
                        w.push((esim, neighbor))
                        if len(w) > ef:
                            w.pop()
        res = RustHeap()
        for pair in w.into_vec():                                                  # core.rs:670-673: raw array order
            res.push(pair)
        return res

    # ---- core.rs:677-757 (extend_candidates = keep_pruned_connections = true at every call site)
    def select_neighbors(self, query, c, m, lc, ignored):
        r = RustHeap()
        w = c.clone()
        wd = RustHeap()
        ccopy = c.clone()
        v = set()
        while len(ccopy):
            v.add(ccopy.pop()[1])
        ccopy = c.clone()
        while len(ccopy):
            _, e = ccopy.pop()
            for en in self.nodes[e].neighbors[lc]:
                if en == query or (ignored is not None and en == ignored):
                    continue
                if en not in v:
                    w.push((self.mfunc(self.nodes[query].data, self.nodes[en].data), en))
                    v.add(en)
        while len(w) and len(r) < m:                                               # core.rs:724-738
            epair = w.pop()
            if epair[1] == query or (ignored is not None and epair[1] == ignored):
                continue
            if len(r) == 0 or epair[0] > r.peek()[0]:
                r.push(epair)
            else:
                wd.push(epair)
        while len(wd) and len(r) < m:                                              # core.rs:741-754
            ppair = wd.pop()
            if ppair[1] == query or (ignored is not None and ppair[1] == ignored):
                continue
            r.push(ppair)
        return r

    # ---- core.rs:759-774
    def connect_neighbors(self, query, neighbors, level):
        nb = neighbors.clone()
        while len(nb):
            _, n = nb.pop()
            self.nodes[query].add_neighbor(level, n)
            self.nodes[n].add_neighbor(level, query)

    # ---- core.rs:776-822
    def update_node_connections(self, node, new_neighbors, old_neighbors, level, ignored):
        newconn = new_neighbors.clone()
        rmconn = old_neighbors.clone().into_vec()
        updated = {node}
        while len(newconn):
            _, n = newconn.pop()
            self.nodes[node].add_neighbor(level, n)
            self.nodes[n].add_neighbor(level, node)
            updated.add(n)
            for i, (_, x) in enumerate(rmconn):                                    # position() + remove()
                if x == n:
                    del rmconn[i]
                    break
        while rmconn:
            _, x = rmconn.pop()                                                    # Vec::pop: from the back
            self.nodes[node].rm_neighbor(level, x)
            if ignored is not None and x == ignored:
                continue
            self.nodes[x].rm_neighbor(level, node)
            updated.add(x)
        return updated

    # ---- core.rs:414-475
    def delete_node(self, nid):
        node = self.nodes.pop(nid)
        self.node_count -= 1
        for lc in range(self.max_layer, -1, -1):
            if lc < len(self.layers) and nid in self.layers[lc]:
                self.layers[lc].remove(nid)
                break
        self.nodes[nid] = node                  # the Arc is still alive while its neighbours are repaired
        updated = set()
        for lc in range(len(node.neighbors)):
            updated |= self.delete_node_from_neighbors(nid, lc)
        del self.nodes[nid]
        self.levels[nid] = -1
        self.last_updated = updated
        if self.enterpoint == nid:              # core.rs:449-472
            new_ep = None
            for lc in range(self.max_layer, -1, -1):
                if self.layers[lc]:
                    new_ep = min(self.layers[lc])   # the reference takes an arbitrary HashSet element (core.rs:453)
                    break
                self.layers.pop()
                if self.max_layer > 0:
                    self.max_layer -= 1
            self.enterpoint = new_ep

    # ---- core.rs:824-863
    def delete_node_from_neighbors(self, nid, lc):
        updated = set()
        for n in list(self.nodes[nid].neighbors[lc]):   # the victim's own list is never edited while walked (:810-813)
            nconn = RustHeap()
            for nn in self.nodes[n].neighbors[lc]:
                nconn.push((self.mfunc(self.nodes[n].data, self.nodes[nn].data), nn))
            m_max = self.m_max_0 if lc == 0 else self.m_max
            nnewconn = self.select_neighbors(n, nconn, m_max, lc, nid)
            updated.add(n)
            updated |= self.update_node_connections(n, nnewconn, nconn, lc, nid)
        return updated

    # ---- core.rs:477-486, 865-892
    def search_knn(self, query, k, ef=None):
        assert len(query) == self.data_dim
        if self.enterpoint is None or self.node_count == 0:
            return [], []
        query = np.asarray(query, dtype=F32)
        ef = self.ef_construction if ef is None else ef
        ep = self.enterpoint
        lc = self.max_layer
        while lc > 0:
            w = self.search_level(query, ep, 1, lc)
            ep = w.peek()[1]
            lc -= 1
        w = self.search_level(query, ep, ef, 0)
        ids, sims = [], []
        while len(ids) < k and len(w):
            s, n = w.pop()
            ids.append(n)
            sims.append(s)
        return ids, sims

    # ---- views for the cross-check
    def adjacency(self, nid, level):
        nb = self.nodes[nid].neighbors
        return list(nb[level]) if level < len(nb) else []
