"""GPU parity of the class-local duplicate detection (csrc/class_dedup.cu) that replaced the global record
sort of large products: against the CPU oracle (symmer/operators/base.py:764-794 + utils.py:230-279 restated),
against the record-sort path it replaces (tuning knob 10 = 0), and on the shapes that stress it — classes
that overflow a CTA, operands from a low-dimensional span (SURVEY.md section 8d's high-collision variant),
forced hash collisions, block lists."""
import numpy as np
import pytest
import torch

from oracle import pauli_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import symmer_b200.ops as o
    o.device()
    return o


def dev_op(ops, symp, coeff):
    n = symp.shape[1] // 2
    return ops.pack(torch.from_numpy(np.ascontiguousarray(symp)), n), torch.from_numpy(np.asarray(coeff, dtype=complex)).cuda()


def host_op(ops, xz, c, n):
    return ops.unpack(xz, n).cpu().numpy(), c.cpu().numpy()


def span_operator(gens, n_rows, rng):
    """n_rows random GF(2) combinations of the generator rows, random complex coefficients."""
    pick = rng.random((n_rows, gens.shape[0])) < 0.5
    symp = (pick.astype(np.uint8) @ gens.astype(np.uint8)) % 2
    return symp.astype(bool), rng.standard_normal(n_rows) + 1j * rng.standard_normal(n_rows)


def check_both_paths(ops, a_s, a_c, b_s, b_c, thr=1e-15, scale_mult=1.0, order=True):
    """Tiled product through the class mode and through the record sort: both against the oracle, rows in the
    reference's first-occurrence order, and against each other bit for bit (rows) / 1e-12 (coefficients)."""
    n = a_s.shape[1] // 2
    ref_s, ref_c = po.multiply_by_operator(a_s, a_c, b_s, b_c, thr)
    scale = max(1e-300, np.abs(a_c).max() * np.abs(b_c).max()) * scale_mult
    out = {}
    try:
        ops.set_tuning(0, 0)                      # force the large-product (ordered-tile) path
        # (knob 10, knob 11): class mode with the compact kernel (default), with the 8-byte-entry kernel, and the record sort
        for knob in ((1, 2), (1, 3), (1, 1), (0, 2)):
            ops.set_tuning(10, knob[0])
            ops.set_tuning(11, knob[1])
            xz, c = ops.mul_cleanup(*dev_op(ops, a_s, a_c), *dev_op(ops, b_s, b_c), thr)
            s, cc = host_op(ops, xz, c, n)
            ok, why = po.compare_term_sets(s, cc, ref_s, ref_c, scale=scale)
            assert ok, (knob, why)
            if order and len(cc) == len(ref_c):
                assert np.array_equal(s, ref_s), knob
            out[knob] = (s, cc)
    finally:
        ops.set_tuning(10, 1)
        ops.set_tuning(11, 2)
        ops.set_tuning(0, 1 << 22)
    base = out[(0, 2)]
    for knob in ((1, 2), (1, 3), (1, 1)):
        if len(base[1]) == len(out[knob][1]):
            assert np.array_equal(base[0], out[knob][0]), knob
            assert np.allclose(base[1], out[knob][1], rtol=1e-12, atol=1e-12 * scale), knob
    return out[(1, 2)]


@pytest.mark.parametrize("n,m1,m2", [(1000, 700, 300), (64, 1500, 400), (200, 3000, 150), (30, 2000, 700)])
def test_class_mode_matches_oracle_and_sort_path(ops, n, m1, m2):
    a_s, a_c = po.random_operator(n, m1, seed=7 * n + m1)
    b_s, b_c = po.random_operator(n, m2, seed=7 * n + m2 + 1)
    k = m2 // 3
    b_s[:k] = a_s[:k]                              # identity terms and repeated products
    a_s[m1 // 2:] = a_s[: m1 - m1 // 2]            # duplicated rows inside A: every cross term has a twin
    check_both_paths(ops, a_s, a_c, b_s, b_c)
    check_both_paths(ops, a_s, a_c, b_s, b_c, thr=0.7)     # drops singletons and group sums
    check_both_paths(ops, a_s, a_c, b_s, b_c, thr=None)    # keeps exact zeros


def test_class_overflow_takes_the_global_sort(ops):
    """All rows of A equal: every cross term of a B row lands in one class, far beyond what a CTA holds, so the
    overflow array + global sort + second group pass run; mixed with ordinary classes."""
    n = 200
    a_s, a_c = po.random_operator(n, 6000, seed=3)
    b_s, b_c = po.random_operator(n, 60, seed=4)
    a_s[:5500] = a_s[0]                            # 5500 identical rows: classes of 5500+ records
    s, cc = check_both_paths(ops, a_s, a_c, b_s, b_c, scale_mult=5500)
    assert len(cc) <= 60 * 501
    # everything in ONE class (A = copies of one row, B = copies of another): a single survivor
    a_s[:] = a_s[0]
    b_s[:] = b_s[1]
    s, cc = check_both_paths(ops, a_s, a_c, b_s, b_c, scale_mult=6000 * 60)
    assert len(cc) <= 1


def test_span_operands_high_collision(ops):
    """both operands from the span of a few generators, so the cross terms collide heavily
    (U <= 2^g) and the group pass does real merging — sums in np.add.at order, exact cancellations."""
    rng = np.random.default_rng(11)
    for g, n, m1, m2 in [(8, 1000, 900, 500), (12, 1000, 1000, 400), (20, 128, 2500, 400), (3, 64, 500, 300)]:
        gens = rng.random((g, 2 * n)) < 0.3
        a_s, a_c = span_operator(gens, m1, rng)
        b_s, b_c = span_operator(gens, m2, rng)
        s, cc = check_both_paths(ops, a_s, a_c, b_s, b_c, scale_mult=max(1.0, m1 * m2 / 2.0 ** g) * 4)
        assert len(cc) <= 2 ** g


def test_class_mode_forced_hash_collisions(ops):
    """Hash equality is only a filter: with the hash masked down to a few bits every record has same-hash
    mates that are different rows; results must not change."""
    a_s, a_c = po.random_operator(100, 400, seed=11)
    b_s, b_c = po.random_operator(100, 90, seed=12)
    b_s[:40] = a_s[:40]
    try:
        for mask in (0xFFFF000000000000, 0xFF00000000000000, 0x0):
            ops.set_debug_key_mask(mask)
            check_both_paths(ops, a_s, a_c, b_s, b_c)
    finally:
        ops.set_debug_key_mask(0xFFFFFFFFFFFFFFFF)


def test_class_mode_block_lists(ops):
    """sym_mul_blocks_*: disjoint rectangles; equal rows in different blocks must merge (they share a class)."""
    n, M, N = 1000, 2000, 300
    a_s, a_c = po.random_operator(n, M, seed=31)
    b_s, b_c = po.random_operator(n, N, seed=32)
    b_s[:100] = a_s[:100]
    b_s[100:200] = a_s[1000:1100]
    a, ac = dev_op(ops, a_s, a_c)
    b, bc = dev_op(ops, b_s, b_c)
    for blocks in ([(0, 900, 150, 300), (900, 901, 0, 150), (901, 2000, 0, 120), (0, 700, 0, 150)],
                   [(int(p), int(p) + 125, 0 if i % 2 else 150, 150 if i % 2 else 300) for i, p in enumerate(range(0, 2000, 125))],
                   [(int(p), int(p) + 100, 0, 300) for p in range(0, 2000, 100)]):        # 4, 16 and 20 blocks (20: sort path)
        rows, coeffs = [], []
        for p0, p1, q0, q1 in blocks:
            r, c = po.cross_terms(a_s[p0:p1], a_c[p0:p1], b_s[q0:q1], b_c[q0:q1])
            rows.append(r)
            coeffs.append(c)
        ref_s, ref_c = po.symplectic_cleanup(np.vstack(rows), np.hstack(coeffs), 1e-15)
        xz, c, T = ops.mul_blocks_cleanup(a, ac, b, bc, blocks)
        s, cc = host_op(ops, xz, c, n)
        ok, why = po.compare_term_sets(s, cc, ref_s, ref_c, scale=np.abs(a_c).max() * np.abs(b_c).max())
        assert ok, (len(blocks), why)
        if len(cc) == len(ref_c):
            assert np.array_equal(s, ref_s)            # block-by-block first-occurrence order


def test_class_mode_many_classes_medium_size(ops):
    """5e6 cross terms at 1000 qubits (thousands of classes, every CTA loops over several) with planted
    duplicates: the survivor count is known exactly and every output row is checked against the XOR of its
    operands; the planted groups against the oracle on the rows involved."""
    n, M, N = 1000, 5000, 1000
    a_s, a_c = po.random_operator(n, M, seed=41)
    b_s, b_c = po.random_operator(n, N, seed=42)
    # plant, for i < 200:  A[i] ^ B[i] == A[i + 2500] ^ B[i + 500]  (and therefore also A[i + 2500] ^ B[i] == A[i] ^ B[i + 500])
    for i in range(200):
        b_s[i + 500] = a_s[i] ^ b_s[i] ^ a_s[i + 2500]
    a, ac = dev_op(ops, a_s, a_c)
    b, bc = dev_op(ops, b_s, b_c)
    xz, c = ops.mul_cleanup(a, ac, b, bc)
    assert xz.shape[0] == M * N - 400
    # survivors in t = q*M + p order: the later member of every planted pair is merged into the earlier one
    t_all = torch.arange(M * N, device=a.device)
    dropped = torch.tensor([(i + 500) * M + i + 2500 for i in range(200)] + [(i + 500) * M + i for i in range(200)],
                           device=a.device)
    keep = torch.ones(M * N, dtype=torch.bool, device=a.device)
    keep[dropped] = False
    t_kept = t_all[keep]
    assert torch.equal(xz, a[t_kept % M] ^ b[t_kept // M])
    # coefficients of the merged heads against the oracle
    for i in (0, 57, 199):
        rows, coeffs = po.cross_terms(a_s[[i, i + 2500]], a_c[[i, i + 2500]], b_s[[i, i + 500]], b_c[[i, i + 500]])
        ref_s, ref_c = po.symplectic_cleanup(rows, coeffs, 1e-15)
        t_head = i * M + i
        slot = int((t_kept == t_head).nonzero()[0, 0])
        j = int(np.flatnonzero((ref_s == ops.unpack(xz[slot:slot + 1].contiguous(), n).cpu().numpy()[0]).all(axis=1))[0])
        assert np.isclose(c[slot].item(), ref_c[j], rtol=1e-12, atol=1e-14)


def test_owner_partition_on_structured_operators(ops, hamiltonians):
    """The exchange-free multi-GPU product (GF(2)-linear owner classes) on operators that are NOT random: a molecular
    Hamiltonian squared (rows in a 28-dimensional space, heavy multiplicities) and operands from a 10-generator span.
    Every "rank" is played in turn on one device: the parts must be disjoint, carry their owner class, cover every cross
    term exactly once, and their union must be the oracle product; the linear owner must not collapse onto few ranks."""
    from symmer_b200 import dist as sdist
    h_s, h_c, _ = hamiltonians("H2O_STO3G")
    rng = np.random.default_rng(5)
    gens = rng.random((10, 2 * 200)) < 0.3
    sa, sac = span_operator(gens, 700, rng)
    sb, sbc = span_operator(gens, 300, rng)
    for (a_s, a_c, b_s, b_c, scale_mult) in ((h_s, h_c, h_s, h_c, 1086.0), (sa, sac, sb, sbc, 700 * 300 / 1024.0 * 4)):
        n = a_s.shape[1] // 2
        ref_s, ref_c = po.multiply_by_operator(a_s, a_c, b_s, b_c)
        a, ac = dev_op(ops, a_s, a_c)
        b, bc = dev_op(ops, b_s, b_c)
        for log2g in (1, 3):
            rows, coeffs, generated, sizes = [], [], 0, []
            for r in range(1 << log2g):
                xz, c, info = sdist.owned_product(a, ac, b, bc, log2g, r)
                generated += info["cross_terms_generated"]
                sizes.append(info["cross_terms_generated"])
                if xz.shape[0]:
                    assert bool((ops.owner_classes(xz, log2g) == r).all())
                s, cc = host_op(ops, xz, c, n)
                rows.append(s)
                coeffs.append(cc)
            assert generated == a_s.shape[0] * b_s.shape[0]
            assert max(sizes) <= 2.0 * generated / len(sizes), sizes          # no rank gets more than twice its share
            s, cc = np.vstack(rows), np.hstack(coeffs)
            assert len(np.unique(s, axis=0)) == len(s)
            scale = float(np.abs(a_c).max() * np.abs(b_c).max()) * scale_mult
            ok, why = po.compare_term_sets(s, cc, ref_s, ref_c, scale=scale)
            assert ok, (n, log2g, why)


def test_fused_rotation_large_uses_privatised_tables_and_matches_two_step(ops):
    """A general rotation of 6e5 rows as ONE block-list product (rotate_dedup: the class tables are built with the
    per-CTA privatised counters because A has hundreds of rows per class) against the two-step form (rotate, then
    cleanup) that the small-size tests pin to the oracle; plus the oracle itself on the rows that involve a sample."""
    import math
    n, M = 128, 600_000
    rng = np.random.default_rng(9)
    base_s, _ = po.random_operator(n, 200_000, seed=17)
    idx = rng.integers(0, 200_000, size=M)
    s = base_s[idx]                                                      # every row ~3 times: the dedup merges
    c = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    q_s, _ = po.random_operator(n, 1, seed=18)
    xz, cc = dev_op(ops, s, c)
    q = ops.pack(torch.from_numpy(q_s), n)
    ang = 0.37
    f_xz, f_c = ops.rotate_dedup(xz, cc, q, math.cos(ang), math.sin(ang))
    t_xz, t_c = ops.cleanup(*ops.rotate(xz, cc, q, math.cos(ang), math.sin(ang), 0))
    assert f_xz.shape == t_xz.shape
    pf, pt = ops.lex_order(f_xz).to(torch.int64), ops.lex_order(t_xz).to(torch.int64)
    assert torch.equal(f_xz[pf], t_xz[pt])
    assert torch.allclose(f_c[pf], t_c[pt], rtol=1e-12, atol=1e-12)
    # oracle on a slice: all copies of 300 distinct base rows
    pick = np.isin(idx, np.arange(300))
    ref_s, ref_c = po.perform_rotations(s[pick], c[pick], [(q_s[0], ang)])
    got_s, got_c = host_op(ops, f_xz, f_c, n)
    keys = {r.tobytes(): v for r, v in zip(got_s, got_c)}
    for r, v in zip(ref_s, ref_c):
        assert np.isclose(keys[r.tobytes()], v, rtol=1e-12, atol=1e-12)
