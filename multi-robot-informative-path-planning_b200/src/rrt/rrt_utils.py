"""Tree data model and RRT primitives of the drop-in (API of the reference's src/rrt/rrt_utils.py).

B200-first layout: a tree is a structure of arrays (`TreeArrays`: coordinates, insertion-time parent,
current parent, a time-stamped rewire log) that uploads to the GPU as-is and is what the branch-gain
kernels walk (ipp_branch_gain).  Node OBJECTS with the reference's attributes (`point`, `parent`,
`children`, `information`, `cost`) are materialised from the arrays only for API compatibility
(`strategy.tree_nodes`, `get_current_node`, plotting hooks).

Decisions are kept bit-identical to the reference:
  * candidate filtering is vectorised, but every comparison that can flip on a rounding (nearest
    ties, the `< d` radius test, path-cost comparisons) is re-evaluated on the short candidate list
    with numpy's own scalar `np.linalg.norm`, exactly the expression of rrt_utils.py:71-101;
  * `add_node` makes the NEAREST node the parent whatever `choose_parent` picked (rrt_utils.py:26-29,
    103-105); `rewire` edits the parent pointer only, never `children` (rrt_utils.py:85-88).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

_norm = np.linalg.norm


# ---- reference-shaped node objects ----------------------------------------------------------------
class TreeNode:
    """rrt_utils.py:10-29."""

    def __init__(self, point: np.ndarray, parent: Optional["TreeNode"] = None, cost: float = 0):
        self.point = point
        self.parent = parent
        self.children: List["TreeNode"] = []
        self.cost = cost

    def add_child(self, child: "TreeNode") -> None:
        self.children.append(child)
        child.parent = self


class InformativeTreeNode(TreeNode):
    """rrt_utils.py:31-35."""

    def __init__(self, point: np.ndarray, parent: Optional[TreeNode] = None):
        super().__init__(point, parent)
        self.information = 0


class TreeCollection:
    """rrt_utils.py:37-54."""

    def __init__(self):
        self.trees: List[TreeNode] = []

    def add(self, tree: TreeNode) -> None:
        self.trees.append(tree)

    def add_repeated(self, tree: TreeNode, count: int) -> None:
        """`count` calls of add(tree): the reference's generators register the root once per loop iteration
        (rrt.py:278,327) -- 18 000 times per tree at the size of its example config."""
        if count > 0:
            self.trees.extend([tree] * count)

    def __iter__(self):
        return iter(self.trees)

    def __getitem__(self, idx: int) -> TreeNode:
        return self.trees[idx]

    def __len__(self) -> int:
        return len(self.trees)


# ---- object-level primitives (same names / semantics as the reference, for external callers) ------
def node_selection_key_distance(node: TreeNode, target_point: np.ndarray) -> float:
    return _norm(node.point - target_point)


def cost(node: TreeNode) -> float:
    total = 0
    while node.parent is not None:
        total += _norm(node.point - node.parent.point)
        node = node.parent
    return total


def line_cost(x1: TreeNode, x2: TreeNode) -> float:
    return _norm(x1.point - x2.point)


def obstacle_free(x1: np.ndarray, x2: np.ndarray) -> bool:
    return True  # rrt_utils.py:81-83: the reference never rejects an edge


def choose_parent(X_near: List[TreeNode], x_nearest: TreeNode, x_new: TreeNode) -> TreeNode:
    best, best_cost = x_nearest, cost(x_nearest) + line_cost(x_nearest, x_new)
    for cand in X_near:
        c = cost(cand) + line_cost(cand, x_new)
        if obstacle_free(cand.point, x_new.point) and c < best_cost:
            best, best_cost = cand, c
    x_new.parent = best
    return x_new


def rewire(X_near: List[TreeNode], x_new: TreeNode) -> None:
    for cand in X_near:
        if obstacle_free(x_new.point, cand.point) and cost(x_new) + line_cost(x_new, cand) < cost(cand):
            cand.parent = x_new


def near(x_new: TreeNode, tree_nodes: List[TreeNode], d_waypoint_distance: float) -> List[TreeNode]:
    return [nd for nd in tree_nodes if _norm(x_new.point - nd.point) < d_waypoint_distance]


def steer(x_nearest: TreeNode, x_rand: np.ndarray, d_max_step: float) -> np.ndarray:
    return steer_point(x_nearest.point, x_rand, d_max_step)


def nearest(tree_nodes: List[TreeNode], x_rand: np.ndarray) -> TreeNode:
    return min(tree_nodes, key=lambda nd: _norm(nd.point - x_rand))


def add_node(tree_nodes: List[TreeNode], x_new: TreeNode, x_parent: TreeNode) -> None:
    x_parent.add_child(x_new)
    tree_nodes.append(x_new)


def trace_path_to_root(selected_leaf: TreeNode) -> List[np.ndarray]:
    out = []
    node = selected_leaf
    while node is not None:
        out.append(node.point)
        node = node.parent
    out.reverse()
    return out


def steer_point(origin: np.ndarray, target: np.ndarray, d_max_step: float) -> np.ndarray:
    """rrt_utils.py:93-98, operation for operation (the coordinates feed every later decision)."""
    direction = target - origin
    distance = _norm(direction)
    direction = direction / distance if distance > 0 else direction
    distance = min(distance, d_max_step)
    return origin + direction * distance


# ---- native tree growth (csrc/ipp_tree_host.cu) --------------------------------------------------------
NATIVE_GROWTH = True          # False: always the Python loops of rrt.py (same results; the tests compare the two)
_NATIVE = {"checked": False, "variant": None, "lib": None}


def native_norm_variant() -> Optional[int]:
    """Which rounding form of ipp_host_norm2 equals np.linalg.norm on this machine (numpy's norm of a 1-D vector is
    sqrt(x.dot(x)); whether that dot fuses its second multiply-add depends on the BLAS numpy links).  Calibrated
    once against np.linalg.norm itself on 4096 seeded vectors; None when no form reproduces all of them, in which
    case the growth stays on the Python loops."""
    if _NATIVE["checked"]:
        return _NATIVE["variant"]
    from mripp_b200 import _lib
    try:
        lib = _lib.load()
    except _lib.IppError:
        # no built library: the Python loops give the same trees (the device paths still refuse to run without it)
        _NATIVE.update(checked=True, variant=None)
        return None
    rng = np.random.RandomState(20240711)
    xy = np.ascontiguousarray(rng.standard_normal((4096, 2)) * 10.0 ** rng.uniform(-3, 3, (4096, 1)))
    xy[:64] = rng.uniform(0, 100, (64, 2)) - rng.uniform(0, 100, (64, 2))     # the coordinate differences of a run
    want = np.array([np.linalg.norm(v) for v in xy])
    out = np.empty(len(xy))
    match = []
    try:  # variants 1 and 2 execute hardware FMA instructions: only try them where the CPU has them
        has_fma = " fma " in (" " + open("/proc/cpuinfo").read().split("flags", 1)[1].split("\n", 1)[0] + " ")
    except (OSError, IndexError):
        has_fma = False
    for variant in (range(3) if has_fma else range(1)):
        _lib.check(lib.ipp_host_norm2(variant, xy.ctypes.data, len(xy), out.ctypes.data), "ipp_host_norm2")
        if np.array_equal(out, want):
            match.append(variant)
    _NATIVE.update(checked=True, variant=match[0] if match else None, lib=lib)
    return _NATIVE["variant"]


class _RngMark:
    """Snapshot / restore of the global np.random stream around a block of look-ahead draws.  get_state/set_state
    cost ~85 us each (they rebuild and validate the 624-word key); the MT19937 bit generator also exposes the address
    of its raw state (BitGenerator.ctypes.state_address: uint32 key[624], int pos), which copies in under 1 us.  The
    raw route is used only after it has reproduced a stream once on this numpy (self-test below); uniform draws do
    not touch the legacy Gaussian cache, which lives outside that struct."""
    _raw = None   # None: untested, False: use get_state / set_state, else the byte count

    @classmethod
    def _address(cls):
        bg = np.random.mtrand._rand._bit_generator
        return bg.ctypes.state_address if type(bg).__name__ == "MT19937" else None

    @classmethod
    def _selftest(cls):
        import ctypes as C
        try:
            addr, size = cls._address(), 624 * 4 + 4
            if addr is None:
                return False
            saved = np.random.get_state()
            np.random.random_sample(700)                      # put pos somewhere inside the key
            mark = C.string_at(addr, size)
            a = np.random.random_sample(1500)                 # crosses a regeneration of the key
            after = np.random.get_state()
            C.memmove(addr, mark, size)
            b = np.random.random_sample(1500)
            now = np.random.get_state()
            ok = bool(np.array_equal(a, b) and np.array_equal(after[1], now[1]) and after[2:] == now[2:])
            np.random.set_state(saved)
            return size if ok else False
        except Exception:
            return False

    def __init__(self):
        import ctypes as C
        if _RngMark._raw is None:
            _RngMark._raw = self._selftest()
        addr = self._address() if _RngMark._raw else None
        self._addr = addr
        self._mark = C.string_at(addr, _RngMark._raw) if addr is not None else np.random.get_state()

    def restore(self):
        import ctypes as C
        if self._addr is not None:
            C.memmove(self._addr, self._mark, _RngMark._raw)
        else:
            np.random.set_state(self._mark)


# ---- structure-of-arrays tree ------------------------------------------------------------------------
class RewireLog:
    """The rewire events (time, node, new parent) of a tree in time order: behaves like the list of tuples it used to
    be (len, iteration, indexing, equality with a list), but the events a native growth call reports stay in the int64
    array they arrive in -- a tree of the reference's example config logs ~28 000 of them per call, and turning them
    into tuples and back into an array for the scorer's CSR cost more than a third of the device scoring pass."""
    __slots__ = ("_parts", "_n", "_cache")

    def __init__(self, events=()):
        self._parts, self._n, self._cache = [], 0, None
        for ev in events:
            self.append(ev)

    def append(self, ev) -> None:
        ev = (int(ev[0]), int(ev[1]), int(ev[2]))
        if self._parts and isinstance(self._parts[-1], list):
            self._parts[-1].append(ev)
        else:
            self._parts.append([ev])
        self._n += 1
        self._cache = None

    def extend_array(self, events: np.ndarray) -> None:
        if len(events):
            self._parts.append(np.array(events, dtype=np.int64).reshape(-1, 3))
            self._n += len(events)
            self._cache = None

    def as_array(self) -> np.ndarray:
        """(E, 3) int64, time order."""
        if self._cache is None:
            parts = [np.asarray(q, dtype=np.int64).reshape(-1, 3) for q in self._parts]
            self._cache = np.concatenate(parts) if parts else np.zeros((0, 3), dtype=np.int64)
            self._parts = [self._cache] if parts else []
        return self._cache

    def __len__(self) -> int:
        return self._n

    def __bool__(self) -> bool:
        return self._n > 0

    def __iter__(self):
        return iter([tuple(r) for r in self.as_array().tolist()])

    def __getitem__(self, k):
        r = self.as_array()[k]
        return tuple(r.tolist()) if r.ndim == 1 else [tuple(x) for x in r.tolist()]

    def __eq__(self, other):
        if isinstance(other, RewireLog):
            return bool(np.array_equal(self.as_array(), other.as_array()))
        if isinstance(other, (list, tuple)):
            return list(self) == [tuple(int(v) for v in ev) for ev in other]
        return NotImplemented

    def __repr__(self) -> str:
        return f"RewireLog({list(self)!r})"


class TreeArrays:
    """One agent's tree as flat arrays; node id = insertion order (index into tree_nodes[agent])."""

    def __init__(self, root_point: np.ndarray, capacity: int = 256):
        self._xy = np.empty((capacity, 2), dtype=np.float64)
        self.parent0 = np.full(capacity, -1, dtype=np.int64)  # parent at insertion (defines `children`)
        self.parent = np.full(capacity, -1, dtype=np.int64)   # current parent (after rewires)
        self.edge = np.zeros(capacity, dtype=np.float64)      # |point - current parent|, numpy-exact
        self.nchild = np.zeros(capacity, dtype=np.int64)
        self.info = np.zeros(capacity, dtype=np.float64)
        self._pt: List[Optional[np.ndarray]] = []             # the original point objects, in order (None: a natively grown
                                                              # node whose array is made from _xy on first use, see _p)
        self.rewires = RewireLog()                            # (time, node, new_parent)
        self.n = 0
        self._append(root_point, -1)

    # -- storage ----------------------------------------------------------------------------------
    def _grow(self):
        cap = self._xy.shape[0] * 2
        for name in ("_xy", "parent0", "parent", "edge", "nchild", "info"):
            old = getattr(self, name)
            new = np.empty((cap,) + old.shape[1:], dtype=old.dtype)
            new[: old.shape[0]] = old
            if name in ("parent0", "parent"):
                new[old.shape[0]:] = -1
            elif name != "_xy":
                new[old.shape[0]:] = 0
            setattr(self, name, new)

    def _append(self, point, parent):
        if self.n == self._xy.shape[0]:
            self._grow()
        i = self.n
        self._xy[i] = point
        self._pt.append(point)
        self.parent0[i] = self.parent[i] = parent
        self.nchild[i] = 0
        self.info[i] = 0.0
        if parent >= 0:
            self.nchild[parent] += 1
            self.edge[i] = _norm(point - self._p(parent))
        self.n += 1
        return i

    def _p(self, i: int) -> np.ndarray:
        """The point object of node i; nodes inserted by ipp_host_tree_grow get theirs on first use (18 000 small arrays
        per tree of the reference's example config, of which a selector touches a few dozen)."""
        q = self._pt[i]
        if q is None:
            q = self._pt[i] = self._xy[i].copy()
        return q

    def grow_native(self, mode: int, workspace, step: float, radius: float, budget_portion: float, block: int = 256
                    ) -> Optional[int]:
        """One call of rrt / rrt_star / rig tree generation (mode 0 / 1 / 2; rrt.py:263-327) through
        ipp_host_tree_grow: samples come from blocks of pre-drawn uniforms -- random_sample(2k) is the stream of k
        calls of np.random.rand(2) -- and the global RNG is left exactly where the reference's loop leaves it.
        Returns the number of samples drawn (= loop iterations of the reference), or None (nothing done) when the
        native norm could not be calibrated or NATIVE_GROWTH is off."""
        variant = native_norm_variant() if NATIVE_GROWTH else None
        if variant is None:
            return None
        from mripp_b200.exact import RNG_SECTION
        lib = _NATIVE["lib"]
        with RNG_SECTION:
            return self._grow_native_locked(lib, mode, variant, workspace, step, radius, budget_portion, block)

    def _grow_native_locked(self, lib, mode, variant, workspace, step, radius, budget_portion, block) -> int:
        import ctypes as C
        samples = 0
        ws4 = np.array([workspace[0], workspace[1], workspace[2], workspace[3]], dtype=np.float64)
        travelled = C.c_double(0.0)
        consumed, n_out, finished, n_rw = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        while True:
            while self._xy.shape[0] < self.n + block:
                self._grow()
            cap_rw = 4 * (self.n + block) + 1024
            rewires = np.empty((cap_rw, 3), dtype=np.int64)
            scratch = np.empty(self._xy.shape[0], dtype=np.int32)
            mark = _RngMark()
            u = np.random.random_sample(2 * block)
            n_rw.value = 0
            rc = lib.ipp_host_tree_grow(mode, variant, self._xy.ctypes.data, self.parent0.ctypes.data, self.parent.ctypes.data,
                                        self.edge.ctypes.data, self.nchild.ctypes.data, self.n, self._xy.shape[0],
                                        u.ctypes.data, block, ws4.ctypes.data, float(step), float(radius),
                                        float(budget_portion), C.addressof(travelled), rewires.ctypes.data, cap_rw,
                                        C.addressof(n_rw), scratch.ctypes.data, C.addressof(consumed), C.addressof(n_out),
                                        C.addressof(finished))
            if rc != 0:
                raise RuntimeError(f"ipp_host_tree_grow failed with status {rc}")
            samples += consumed.value
            if consumed.value < block:                      # leave the RNG where consumed.value calls of rand(2) leave it
                mark.restore()
                if consumed.value:
                    np.random.random_sample(2 * consumed.value)
            self._pt.extend([None] * (n_out.value - self.n))
            self.info[self.n: n_out.value] = 0.0
            self.n = n_out.value
            if n_rw.value:
                if not isinstance(self.rewires, RewireLog):
                    self.rewires = RewireLog(self.rewires)
                self.rewires.extend_array(rewires[: n_rw.value])
            if finished.value:
                return samples
            # the library rebuilds its grid and path costs from the tree on every call: let the blocks grow with the
            # tree (the reference's example config grows 18 000-node trees: 72 blocks of 256 were 72 rebuilds)
            block = min(2 * block, 8192)

    def points(self) -> np.ndarray:
        return self._xy[: self.n]

    def point(self, i) -> np.ndarray:
        return self._p(i)

    # -- primitives, decision-exact ---------------------------------------------------------------
    def nearest(self, target: np.ndarray) -> int:
        """argmin_i |p_i - target| with the reference's first-minimum tie-break (rrt_utils.py:100-101)."""
        d = self._xy[: self.n] - target
        r = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        m = r.min()
        cand = np.nonzero(r <= m * (1.0 + 1e-12) + 1e-300)[0]
        if cand.size == 1:
            return int(cand[0])
        best, best_d = -1, None
        for i in cand:  # ascending index, strict '<' keeps the first minimum
            di = _norm(self._p(i) - target)
            if best_d is None or di < best_d:
                best, best_d = int(i), di
        return best

    def near(self, point: np.ndarray, radius: float) -> List[int]:
        """[i : |point - p_i| < radius], ascending (rrt_utils.py:90-91)."""
        d = self._xy[: self.n] - point
        r = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        out = []
        for i in np.nonzero(r < radius * (1.0 + 1e-12))[0]:
            if r[i] < radius * (1.0 - 1e-12) or _norm(point - self._p(i)) < radius:
                out.append(int(i))
        return out

    def path_cost(self, i: int) -> float:
        """rrt_utils.py:71-76: leaf-to-root accumulation of the numpy-exact edge lengths."""
        total = 0
        while self.parent[i] >= 0:
            total += self.edge[i]
            i = self.parent[i]
        return total

    def add(self, point: np.ndarray, parent: int) -> int:
        return self._append(point, parent)

    def rewire(self, neighbours: List[int], new: int) -> None:
        """rrt_utils.py:85-88; logs (time = id of the inserted node, node, new parent) for the
        versioned branch walk of ipp_branch_gain."""
        pn = self._p(new)
        for q in neighbours:
            if self.path_cost(new) + _norm(pn - self._p(q)) < self.path_cost(q):
                self.parent[q] = new
                self.edge[q] = _norm(self._p(q) - pn)
                self.rewires.append((new, q, new))

    def history_csr(self):
        """(hist_off[n+1], hist_time[E], hist_parent[E]) int32, events of a node in time order."""
        E = len(self.rewires)
        off = np.zeros(self.n + 1, dtype=np.int32)
        if E == 0:
            return off, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
        ev = self.rewires.as_array() if isinstance(self.rewires, RewireLog) else np.asarray(self.rewires, dtype=np.int64)
        order = np.lexsort((ev[:, 0], ev[:, 1]))  # by node, then time (stable)
        ev = ev[order]
        counts = np.bincount(ev[:, 1], minlength=self.n)
        off[1:] = np.cumsum(counts)
        return off, ev[:, 0].astype(np.int32), ev[:, 2].astype(np.int32)

    def chain_at(self, i: int) -> List[np.ndarray]:
        """Points from node i up to the root following the parents as they were when node i was
        inserted (rewire events with time >= i are ignored) -- the walk ipp_branch_gain does."""
        hist = {}
        for t, q, p in self.rewires:
            if t < i:
                hist[q] = p  # events are appended in time order: the last one before i wins
        out, q = [self._p(i)], i
        while True:
            p = hist.get(q, int(self.parent0[q]))
            if p < 0:
                break
            out.append(self._p(p))
            q = p
        return out

    # -- traversal --------------------------------------------------------------------------------
    def leaves_in_order(self) -> np.ndarray:
        """Nodes without children, in insertion order (rrt.py:374,443)."""
        return np.nonzero(self.nchild[: self.n] == 0)[0]

    def leaves_dfs(self) -> List[int]:
        """Leaves reached from the root through the insertion-time `children` lists, depth first
        (rrt.py:389-395)."""
        if NATIVE_GROWTH and self.n > 256:   # the same traversal natively: a tree of the reference's example config has
            try:                              # ~18 000 nodes, 50 ms per call in the loop below
                from mripp_b200 import _lib
                import ctypes as C
                lib = _lib.load()
                out = np.empty(self.n, dtype=np.int32)
                m = C.c_int(0)
                p0 = np.ascontiguousarray(self.parent0[: self.n], dtype=np.int64)
                if lib.ipp_host_leaves_dfs(p0.ctypes.data, self.n, out.ctypes.data, C.addressof(m)) == 0:
                    return out[: m.value].tolist()
            except Exception:
                pass
        kids: List[List[int]] = [[] for _ in range(self.n)]
        for i in range(1, self.n):
            kids[self.parent0[i]].append(i)
        out, stack = [], [0]
        while stack:
            i = stack.pop()
            if not kids[i]:
                out.append(i)
            stack.extend(reversed(kids[i]))
        return out

    def path_to_root(self, i: int) -> List[np.ndarray]:
        """rrt_utils.py:107-114 over the CURRENT parent pointers; returns the original point objects."""
        out = []
        while i >= 0:
            out.append(self._p(i))
            i = self.parent[i]
        out.reverse()
        return out

    def materialise(self) -> List[InformativeTreeNode]:
        """Reference-shaped node objects (children from insertion-time parents, parent = current)."""
        nodes = [InformativeTreeNode(self._p(i)) for i in range(self.n)]
        for i in range(1, self.n):
            nodes[self.parent0[i]].children.append(nodes[i])
        for i in range(self.n):
            nodes[i].parent = nodes[self.parent[i]] if self.parent[i] >= 0 else None
            nodes[i].information = self.info[i] if i else 0
        return nodes
