"""Seeded synthetic inputs for tests, goldens and ``bench.py`` (NOT part of the product path).

Shapes follow the reference's own generator (``network_simulation/create_test_network.py:11-36``
— random digraph, in-degree <= 3, ~20 % inhibitory edges — and
``network_simulation/network_simulation.py:64-88,245-252`` — cells relaxed under planted rules,
then bit-flip noise) but nothing is shared with it: graphs, rules and cells are drawn from a
``numpy.random.Generator`` and cells are produced directly as 32-bit cell bitplanes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# truth tables of the slots themselves under the lop3 convention bit = (A<<2)|(B<<1)|C
LUT_A, LUT_B, LUT_C = 0xF0, 0xCC, 0xAA


@dataclass
class NodeSpec:
    name: str
    index: int
    predecessors: dict            # ordered {row: name}
    inversions: dict              # ordered {row: bool}
    logic: str = "A"              # planted rule text over A, B, C
    lut: int = LUT_A              # its truth table
    extra: dict = field(default_factory=dict)


def _planted_rule(rng, n_parents: int, inverted) -> tuple:
    """A random read-once AND/OR rule over the bound slots with the PKN inversions applied."""
    lits = [("not " if inv else "") + s for s, inv in zip("ABC", inverted)]
    if n_parents == 1:
        text = lits[0]
    elif n_parents == 2:
        text = f"{lits[0]} {rng.choice(['and', 'or'])} {lits[1]}"
    else:
        form = int(rng.integers(0, 4))
        a, b, c = lits
        text = [f"{a} and {b} and {c}", f"{a} or {b} or {c}",
                f"({a} or {b}) and {c}", f"{a} and {b} or {c}"][form]
    return text, lut_of(text)


def lut_of(logic: str) -> int:
    table = 0
    for idx in range(8):
        env = {"A": bool(idx & 4), "B": bool(idx & 2), "C": bool(idx & 1)}
        if eval(logic, {"__builtins__": {}}, env):
            table |= 1 << idx
    return table


def random_network(n_nodes: int, seed: int, p_inhibit: float = 0.2, p_source: float = 0.1,
                   allow_self: bool = False) -> list:
    """``n_nodes`` NodeSpecs; each non-source node gets 1-3 distinct parents."""
    rng = np.random.default_rng(seed)
    specs = []
    for i in range(n_nodes):
        name = f"G{i}"
        if n_nodes == 1 or rng.random() < p_source:
            specs.append(NodeSpec(name, i, {}, {}, "A", LUT_A))
            continue
        k = int(rng.integers(1, 4))
        pool = [j for j in range(n_nodes) if allow_self or j != i]
        k = min(k, len(pool))
        parents = [int(x) for x in rng.choice(pool, size=k, replace=False)]
        inverted = [bool(rng.random() < p_inhibit) for _ in parents]
        logic, lut = _planted_rule(rng, k, inverted)
        specs.append(NodeSpec(name, i, {p: f"G{p}" for p in parents},
                              {p: v for p, v in zip(parents, inverted)}, logic, lut))
    return specs


def spec_parent_rows(spec: NodeSpec) -> tuple:
    keys = list(spec.predecessors)[:3] if spec.predecessors else [spec.index]
    return tuple(keys + [-1] * (3 - len(keys)))


def eval_lut_words(lut: int, a, b, c):
    """Truth-table evaluation on packed words (any unsigned NumPy dtype)."""
    out = np.zeros_like(a)
    full = ~np.zeros_like(a)
    for idx in range(8):
        if (lut >> idx) & 1:
            term = (a if idx & 4 else a ^ full) & (b if idx & 2 else b ^ full) & \
                   (c if idx & 1 else c ^ full)
            out |= term
    return out


def random_planes(n_rows: int, n_cells: int, seed: int, density=None) -> np.ndarray:
    """uint32 bitplanes [n_rows][ceil(n_cells/32)], bit b of word w = cell 32*w+b, tail zero."""
    rng = np.random.default_rng(seed)
    n_words = (n_cells + 31) // 32
    planes = np.empty((n_rows, n_words), dtype=np.uint32)
    if density is None:
        density = rng.uniform(0.1, 0.5, size=n_rows)
    density = np.broadcast_to(np.asarray(density, dtype=np.float64), (n_rows,))
    for r in range(n_rows):
        bits = rng.random(n_words * 32) < density[r]
        bits[n_cells:] = False
        planes[r] = np.packbits(bits, bitorder="little").view(np.uint32)
    return planes


def relax_and_flip(specs, planes: np.ndarray, n_cells: int, seed: int, relax_steps: int = 30,
                   flip: float = 0.05) -> np.ndarray:
    """Run ``relax_steps`` synchronous updates under the planted rules, then flip bits."""
    rng = np.random.default_rng(seed + 7919)
    state = planes.copy()
    zeros = np.zeros_like(state[0])
    for _ in range(relax_steps):
        nxt = np.empty_like(state)
        for s in specs:
            ra, rb, rc = spec_parent_rows(s)
            nxt[s.index] = eval_lut_words(s.lut, state[ra], state[rb] if rb >= 0 else zeros,
                                          state[rc] if rc >= 0 else zeros)
        state = nxt
    if flip > 0:
        n_words = state.shape[1]
        for r in range(state.shape[0]):
            noise = rng.random(n_words * 32) < flip
            state[r] ^= np.packbits(noise, bitorder="little").view(np.uint32)
    mask_tail(state, n_cells)
    return state


def mask_tail(planes: np.ndarray, n_cells: int) -> None:
    rem = n_cells % 32
    if rem:
        planes[:, n_cells // 32] &= np.uint32((1 << rem) - 1)
    planes[:, (n_cells + 31) // 32:] = 0


def planes_to_dense(planes: np.ndarray, n_cells: int) -> np.ndarray:
    """uint8 (rows, n_cells) of 0/1."""
    bits = np.unpackbits(planes.view(np.uint8), axis=1, bitorder="little")
    return np.ascontiguousarray(bits[:, :n_cells])


def dense_to_planes(dense01: np.ndarray) -> np.ndarray:
    n_rows, n_cells = dense01.shape
    n_words = (n_cells + 31) // 32
    padded = np.zeros((n_rows, n_words * 32), dtype=np.uint8)
    padded[:, :n_cells] = dense01 != 0
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(n_rows, n_words)


def synthetic_dataset(n_nodes: int, n_cells: int, seed: int, flip: float = 0.05,
                      relax_steps: int = 30, **net_kw):
    """(specs, planes) — planted network + relaxed, noised cells."""
    specs = random_network(n_nodes, seed, **net_kw)
    start = random_planes(n_nodes, n_cells, seed + 1)
    planes = relax_and_flip(specs, start, n_cells, seed, relax_steps=relax_steps, flip=flip)
    return specs, planes
