"""Property tests (hypothesis) of the host-side invariants the device paths lean on: static batch shapes, disjoint
splits, shard ownership, metric additivity, the MX-FP8 quantisation rule, wire formats.  CPU only."""
import numpy as np
import pytest
import torch

pytest.importorskip('hypothesis')
from hypothesis import assume, given, settings, strategies as st  # noqa: E402

from coinstac_dinunet_b200.data import COINNPaddedDataSampler
from coinstac_dinunet_b200.data.datautils import create_k_fold_splits, create_ratio_split
from coinstac_dinunet_b200.metrics import COINNAverages, ConfusionMatrix, Prf1a

FAST = settings(max_examples=60, deadline=None, derandomize=True, database=None)   # same examples every run


class _Sized:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


@FAST
@given(n=st.integers(1, 300), bs=st.integers(1, 40), seed=st.integers(0, 2 ** 20), epoch=st.integers(0, 50),
       shuffle=st.booleans())
def test_padded_sampler_yields_whole_batches_that_cover_the_dataset(n, bs, seed, epoch, shuffle):
    """Every site runs ceil(n / batch) FULL batches (static shapes -> CUDA-graph capture), every sample is seen, padding wraps
    around, and the order is a pure function of (seed, epoch) so replicas of a site agree (ref data.py:216-258)."""
    s = COINNPaddedDataSampler(_Sized(n), bs, seed=seed, shuffle=shuffle)
    s.set_epoch(epoch)
    order = list(s)
    assert len(order) == len(s) == -(-n // bs) * bs
    assert set(order) == set(range(n))
    assert order[n:] == order[:len(order) - n]                      # the padding is the head of the same permutation
    s2 = COINNPaddedDataSampler(_Sized(n), bs, seed=seed, shuffle=shuffle)
    s2.set_epoch(epoch)
    assert list(s2) == order
    if not shuffle:
        assert order[:n] == list(range(n))


@FAST
@given(n=st.integers(2, 200), k=st.integers(2, 10))
def test_k_fold_splits_partition_the_files(n, k, tmp_path_factory):
    """Fold i: test = chunk i, validation = chunk i+1, train = the rest; the three parts are disjoint and cover every file,
    and over the k folds every file is tested exactly once (ref datautils.py:44-66)."""
    if k > n:
        k = n
    out = tmp_path_factory.mktemp('folds')
    files = [f'f{i:04d}' for i in range(n)]
    create_k_fold_splits(list(files), {'num_folds': k, 'split_dir': str(out)})
    import json
    tested = []
    for i in range(k):
        fold = json.load(open(out / f'SPLIT_{i}.json'))
        parts = [fold['train'], fold['validation'], fold['test']]
        flat = sum(parts, [])
        if k > 2:
            assert len(flat) == len(set(flat)) == n
        else:                                                        # k == 2: validation chunk == the other fold's test chunk
            assert set(flat) == set(files)
        assert set(flat) == set(files)
        assert not set(fold['test']) & set(fold['train']) and not set(fold['validation']) & set(fold['train'])
        tested += fold['test']
    assert sorted(tested) == files


@FAST
@given(n=st.integers(0, 300), a=st.floats(0.05, 0.9), b=st.floats(0.0, 0.5))
def test_ratio_split_is_a_partition_with_rounding_surplus_in_the_first_key(n, a, b):
    ratio = [a, (1 - a) * b, (1 - a) * (1 - b)]
    files = [f'f{i:04d}' for i in range(n)]
    out = create_ratio_split(list(files), {'split_ratio': ratio}, shuffle_files=True)
    flat = out['train'] + out['validation'] + out['test']
    assert sorted(flat) == files
    assert len(out['test']) == int(ratio[2] * n)
    assert len(out['validation']) == int((ratio[2] + ratio[1]) * n) - int(ratio[2] * n)
    again = create_ratio_split(list(files), {'split_ratio': ratio}, shuffle_files=True)
    assert again == out                                               # seeded by len(files): every site derives the same split


@FAST
@given(world=st.integers(2, 8), sizes=st.lists(st.integers(1, 5000), min_size=1, max_size=6))
def test_shard_ownership_tiles_every_launch_unit_exactly_once(world, sizes):
    """``DistArena.owner_ranges`` (what ``gather_state`` broadcasts): for every sharded launch unit the ranks' ranges are
    disjoint, ordered and cover the unit; one-shot units are replicated and absent."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    units, off = [], 0
    for n in sizes:
        numel = -(-n // (4 * world)) * 4 * world
        units.append((off, numel))
        off += numel
    arena = DistArena.__new__(DistArena)
    arena.world, arena.backend = world, 'nvlink'
    arena._launch_units = lambda: units
    arena._pick_variant = lambda nbytes: 'one_shot' if nbytes <= 4096 else 'two_shot'
    covered = {}
    for q, lo, hi in arena.owner_ranges():
        assert 0 <= q < world and lo < hi and lo % 4 == 0
        for u_off, u_n in units:
            if u_off <= lo < u_off + u_n:
                assert hi <= u_off + u_n
                covered.setdefault(u_off, []).append((lo, hi, q))
    for u_off, u_n in units:
        if u_n * 4 <= 4096:
            assert u_off not in covered
            continue
        spans = sorted(covered[u_off])
        assert spans[0][0] == u_off and spans[-1][1] == u_off + u_n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert [s[2] for s in spans] == sorted(s[2] for s in spans)


@FAST
@given(data=st.data(), parts=st.integers(1, 5))
def test_prf1a_and_confusion_matrix_are_additive_over_batches_and_sites(data, parts):
    """Scores of a concatenation == accumulate() of the pieces (what the per-step device ring relies on).  Across sites both
    classes ship [accuracy, precision, recall] and the aggregate is the unweighted mean of the sites (the reference's
    wire semantics, metrics.py:214-218 / 280-283, remote.py:105-141)."""
    C = data.draw(st.integers(2, 5))
    chunks = []
    for _ in range(parts):
        n = data.draw(st.integers(1, 40))
        pred = torch.tensor(data.draw(st.lists(st.integers(0, C - 1), min_size=n, max_size=n)))
        true = torch.tensor(data.draw(st.lists(st.integers(0, C - 1), min_size=n, max_size=n)))
        chunks.append((pred, true))
    whole_p, whole_t = torch.cat([c[0] for c in chunks]), torch.cat([c[1] for c in chunks])

    cm_all = ConfusionMatrix(num_classes=C)
    cm_all.add(whole_p, whole_t)
    cm_acc, pieces = ConfusionMatrix(num_classes=C), []
    for p, t in chunks:
        m = ConfusionMatrix(num_classes=C)
        m.add(p, t)
        cm_acc.accumulate(m)
        pieces.append(m.serialize())
    assert cm_acc.get() == cm_all.get()
    red = ConfusionMatrix(num_classes=C)
    red.reduce_sites(pieces)
    mean = np.mean(np.asarray(pieces), axis=0)
    assert [red.accuracy(), red.precision(), red.recall()] == pytest.approx(list(mean), abs=1e-9)
    assert int(cm_all.matrix.sum()) == whole_p.numel()

    bp, bt = (whole_p > 0).long(), (whole_t > 0).long()
    f_all = Prf1a()
    f_all.add(bp, bt)
    f_acc, pieces = Prf1a(), []
    for p, t in chunks:
        m = Prf1a()
        m.add((p > 0).long(), (t > 0).long())
        f_acc.accumulate(m)
        pieces.append(m.serialize())
    assert f_acc.get() == f_all.get()
    red = Prf1a()
    red.reduce_sites(pieces)
    mean = np.mean(np.asarray(pieces), axis=0)
    assert [red.accuracy, red.precision, red.recall] == pytest.approx(list(mean), abs=1e-5)
    tp = int(((bp == 1) & (bt == 1)).sum())
    assert f_all.tp == tp and f_all.tp + f_all.fp + f_all.tn + f_all.fn == bp.numel()


@FAST
@given(vals=st.lists(st.tuples(st.floats(-1e3, 1e3), st.integers(1, 64)), min_size=1, max_size=20), split=st.integers(0, 20))
def test_averages_are_count_weighted_and_additive(vals, split):
    whole, a, b = COINNAverages(), COINNAverages(), COINNAverages()
    for i, (v, n) in enumerate(vals):
        whole.add(v, n)
        (a if i < split else b).add(v, n)
    a.accumulate(b)
    total = sum(n for _, n in vals)
    want = sum(v * n for v, n in vals) / total
    assert whole.get()[0] == pytest.approx(want, abs=1e-3)
    assert a.get() == whole.get()
    red = COINNAverages()
    red.reduce_sites([whole.serialize(), whole.serialize()])
    assert red.get()[0] == pytest.approx(want, abs=1e-3)


@FAST
@given(rows=st.integers(1, 6), k=st.integers(1, 300), scale=st.floats(-20, 20), seed=st.integers(0, 1000))
def test_mx_quantisation_rule(rows, k, scale, seed):
    """``quantize_mx_reference`` (the oracle the e4m3 kernels are tested against): per 32-element block the scale is a power of
    two, no element saturates, the block maximum uses the top binade of e4m3 and the round trip is within half an e4m3 ulp."""
    from coinstac_dinunet_b200.ops.fp8 import dequantize_mx, quantize_mx_reference
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, k, generator=g) * (2.0 ** scale)
    q, sf = quantize_mx_reference(x)
    kp = (k + 127) // 128 * 128
    assert q.shape == (rows, kp) and q.dtype == torch.uint8 and sf.shape == (rows, kp // 128)
    back = dequantize_mx(q, sf, k)
    xp = torch.nn.functional.pad(x, (0, kp - k)).view(rows, kp // 32, 32)
    amax = xp.abs().amax(-1, keepdim=True)
    err = (back - x).abs().view(-1)
    # e4m3: 3 mantissa bits -> relative step 2^-3 within a binade; elements far below the block maximum fall into the
    # subnormal range of the scaled format, whose absolute step is 2^-9 x scale <= amax * 2^-9 / 224
    bound = (x.abs() * 2.0 ** -4 + (torch.nn.functional.pad(amax.expand(-1, -1, 32).reshape(rows, kp), (0, 0))[:, :k]) * 2.0 ** -9 / 224 * 1.01
             + 1e-45).view(-1)
    assert bool((err <= bound).all())
    exps = sf.view(torch.uint8).view(rows, kp // 32).float() - 127
    scaled = (amax.squeeze(-1) / torch.exp2(exps))
    live = amax.squeeze(-1) > 2.0 ** -120
    assert bool((scaled[live] <= 448).all()) and bool((scaled[live] > 224 * (1 - 1e-6)).all())


@FAST
@given(shapes=st.lists(st.lists(st.integers(1, 7), min_size=0, max_size=3), min_size=1, max_size=6),
       dtype=st.sampled_from(['float32', 'float16']))
def test_gradient_wire_files_round_trip(shapes, dtype, tmp_path_factory):
    """The reference's wire format: an object array of per-parameter ndarrays in one .npy (tensorutils.py:44-60)."""
    from coinstac_dinunet_b200.utils import tensorutils as tu
    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal(tuple(s)).astype(dtype) for s in shapes]
    path = str(tmp_path_factory.mktemp('wire') / 'grads.npy')
    tu.save_arrays(path, arrays)
    back = tu.load_arrays(path)
    assert len(back) == len(arrays)
    for a, b in zip(arrays, back):
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)


@FAST
@given(n=st.integers(2, 24), m=st.integers(2, 24), rank=st.integers(1, 4), seed=st.integers(0, 10 ** 6))
def test_powersgd_round_invariants(n, m, rank, seed):
    """The PyTorch twin of the batched PowerSGD kernels (``ops.lowrank.powersgd_round_reference``; the GPU tests compare the
    kernels against it): approximation + new error == gradient + old error exactly, the approximation has rank <= r, a
    full-rank budget reproduces the matrix, and with warm start on a fixed matrix the residual falls to the SVD truncation
    error (subspace iteration) - ref powersgd/__init__.py:61-181."""
    from coinstac_dinunet_b200.ops.lowrank import powersgd_round_reference
    assume(rank < min(n, m))                             # PowerSGDPlan sends matrices with min(n, m) <= rank uncompressed
    g = torch.Generator().manual_seed(seed)
    G = torch.randn(n, m, generator=g, dtype=torch.float64)
    E = torch.randn(n, m, generator=g, dtype=torch.float64) * 0.1
    Q = torch.randn(m, rank, generator=g, dtype=torch.float64)
    approx, errs, qs = powersgd_round_reference([G], [E], [Q.clone()], rank, lambda ts: None)
    assert torch.allclose(approx[0] + errs[0], G + E, atol=1e-10)
    assert int(torch.linalg.matrix_rank(approx[0], tol=1e-8)) <= rank
    assert qs[0].shape == (m, rank)

    full = min(n, m)
    Qf = torch.randn(m, full, generator=g, dtype=torch.float64)
    a_full, e_full, _ = powersgd_round_reference([G], [torch.zeros_like(G)], [Qf], full, lambda ts: None)
    if n <= m and float(torch.linalg.cond(G @ Qf)) < 1e4:        # P spans R^n -> P P^T is the identity
        assert float(e_full[0].abs().max()) < 1e-4 * float(G.abs().max())       # (Gram-Schmidt carries an epsilon of 1e-8)

    sv = torch.linalg.svdvals(G)
    best = float(sv[rank:].square().sum().sqrt())
    q = Q.clone()
    for _ in range(60):                                  # warm start: the averaged Q of one step seeds the next
        a, _, qs = powersgd_round_reference([G], [torch.zeros_like(G)], [q], rank, lambda ts: None)
        q = qs[0]
    resid = float((G - a[0]).norm())
    assert resid >= best - 1e-8
    gap = float(sv[rank - 1] - sv[rank]) if rank < len(sv) else 1.0
    if gap > 0.2 * float(sv[0]):                         # well separated spectrum: 60 sweeps converge
        assert resid <= best * 1.001 + 1e-6
