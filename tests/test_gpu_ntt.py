"""GPU parity: NTT through the C ABI vs the reference CPU backend (bit-exact, like icicle/tests/test_mod_arithmetic_api.h:614-695)
and vs the defining sums in Python integers for every NTT field."""
import itertools
import random

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu

NTT_FIELDS = [(ib.Field.BN254_FR, "bn254_fr"), (ib.Field.BLS12_381_FR, "bls12_381_fr"), (ib.Field.BLS12_377_FR, "bls12_377_fr"),
              (ib.Field.BLS12_377_FQ, "bls12_377_fq"), (ib.Field.STARK252, "stark252"), (ib.Field.BABYBEAR, "babybear"),
              (ib.Field.KOALABEAR, "koalabear")]


def omega(name, logn):
    fp = utils.field_params(name)
    return pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])


@pytest.fixture(scope="module")
def domains():
    dom_log = 14
    for field, name in NTT_FIELDS:
        L = utils.field_params(name)["limbs"]
        ib.ntt_release_domain(field)
        ib.ntt_init_domain(field, utils.to_limbs([omega(name, dom_log)], L)[0])
    yield dom_log
    for field, _ in NTT_FIELDS:
        ib.ntt_release_domain(field)


@pytest.mark.parametrize("field,name", NTT_FIELDS)
def test_definition_small(field, name, domains):
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    for logn in (0, 1, 2, 3, 5, 6):
        n = 1 << logn
        w = omega(name, logn)
        x = common.rand_field_elems(name, n, 10 + logn, as_ints=True)
        X = utils.to_limbs(x, L)
        g = 0x1234567 % p
        for inverse, coset in itertools.product((False, True), (1, g)):
            exp = common.ntt_naive_ints(x, w, p, inverse=inverse, coset=coset)
            cfg = ib.NTTConfig(coset_gen=utils.to_limbs([coset], L)[0] if coset != 1 else None)
            got = utils.from_limbs(ib.ntt(field, X, n, ib.NTTDir.kInverse if inverse else ib.NTTDir.kForward, cfg))
            assert got == exp, (name, logn, inverse, coset)
        assert utils.from_limbs(ib.get_root_of_unity_from_domain(field, logn)) == [w]


@pytest.mark.parametrize("field,name", NTT_FIELDS)
def test_roundtrip_and_orderings(field, name, domains):
    fp = utils.field_params(name)
    L = fp["limbs"]
    for logn in (4, 9, 12):
        n = 1 << logn
        X = common.rand_field_elems(name, n, 20 + logn)
        nn = ib.ntt(field, X, n, ib.NTTDir.kForward, ib.NTTConfig(ordering=ib.Ordering.kNN))
        nr = ib.ntt(field, X, n, ib.NTTDir.kForward, ib.NTTConfig(ordering=ib.Ordering.kNR))
        perm = [common.bitrev(i, logn) for i in range(n)]
        assert np.array_equal(nr, nn[perm])
        rn = ib.ntt(field, X[perm], n, ib.NTTDir.kForward, ib.NTTConfig(ordering=ib.Ordering.kRN))
        assert np.array_equal(rn, nn)
        rr = ib.ntt(field, X[perm], n, ib.NTTDir.kForward, ib.NTTConfig(ordering=ib.Ordering.kRR))
        assert np.array_equal(rr, nr)
        back = ib.ntt(field, nn, n, ib.NTTDir.kInverse, ib.NTTConfig())
        assert np.array_equal(back, X)
        # kNM -> kMN round trip (wrappers/rust/icicle-core/src/ntt/tests.rs:159-207)
        nm = ib.ntt(field, X, n, ib.NTTDir.kForward, ib.NTTConfig(ordering=ib.Ordering.kNM))
        mn = ib.ntt(field, nm, n, ib.NTTDir.kInverse, ib.NTTConfig(ordering=ib.Ordering.kMN))
        assert np.array_equal(mn, X)
        # radix-2 algorithm key gives the same answer
        r2 = ib.ntt(field, X, n, ib.NTTDir.kForward, ib.NTTConfig(ext={"ntt_algorithm": 1}))
        assert np.array_equal(r2, nn)


def test_vs_reference_randomized(domains):
    """Random logn / batch / columns / in-place / direction / ordering / coset, bit-exact vs the reference CPU backend."""
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built")
    import torch
    r = ref_icicle.get("bn254")
    name, field = "bn254_fr", ib.Field.BN254_FR
    dom_log = domains
    r.ntt_release_domain()
    r.ntt_init_domain(utils.to_limbs([omega(name, dom_log)], 8)[0])
    rng = random.Random(99)
    fp = utils.field_params(name)
    for trial in range(60):
        logn = rng.randrange(0, 13)
        n = 1 << logn
        batch = 1 << rng.randrange(0, 3)
        columns = rng.random() < 0.5
        inplace = rng.random() < 0.5
        direction = rng.randrange(2)
        ordering = rng.randrange(4)
        coset_kind = rng.randrange(3)
        coset = None
        if coset_kind == 1:  # a domain element: the reference finds its stride in the twiddle table (ntt_data.h:81-83)
            coset = utils.to_limbs([pow(omega(name, dom_log), rng.randrange(1, 3), fp["p"])], 8)[0]
        elif coset_kind == 2:
            coset = utils.to_limbs([rng.randrange(2, fp["p"])], 8)[0]
        x = r.generate_scalars(n * batch)
        exp = r.ntt(x, n, direction, coset_gen=coset, batch_size=batch, columns_batch=columns, ordering=ordering)
        cfg = ib.NTTConfig(coset_gen=coset, batch_size=batch, columns_batch=columns, ordering=ib.Ordering(ordering))
        d_in = ib.to_device(x)
        d_out = d_in if inplace else ib.device_empty(x.size).view(*x.shape)
        ib.ntt(field, d_in, n, direction, cfg, d_out)
        got = ib.to_host(d_out, x.shape)
        assert np.array_equal(got, exp), dict(trial=trial, logn=logn, batch=batch, columns=columns, inplace=inplace, dir=direction, ordering=ordering, coset=coset_kind)
        # host-memory path gives the same
        got_h = ib.ntt(field, x, n, direction, ib.NTTConfig(coset_gen=coset, batch_size=batch, columns_batch=columns, ordering=ib.Ordering(ordering)))
        assert np.array_equal(got_h, exp)
    r.ntt_release_domain()


def test_errors(domains):
    field = ib.Field.BN254_FR
    X = common.rand_field_elems("bn254_fr", 1 << 15, 1)
    with pytest.raises(ib.IcicleError):
        ib.ntt(field, X, 1 << 15, ib.NTTDir.kForward)  # larger than the domain (cpu_ntt_main.h:39-41)
    with pytest.raises(ib.IcicleError):
        ib.ntt(field, X[:24], 24, ib.NTTDir.kForward)  # not a power of two (cpu_ntt_main.h:38)
    with pytest.raises(ib.IcicleError):
        ib.ntt_init_domain(ib.Field.BN254_FQ, X[0])  # field without an NTT in the reference


@pytest.mark.parametrize("field,name,dom_log", [(ib.Field.BABYBEAR, "babybear", 24), (ib.Field.KOALABEAR, "koalabear", 22)])
def test_small_field_tile_pass(field, name, dom_log):
    """The dedicated 4-byte-field pass (csrc/ntt31.cuh, natural order in/out, n >= 2^10) against (a) the oracle port on every
    pass-count / stage-split shape (2^10 .. 2^19: 2 passes of 5..9 stages, 3 passes at 2^19), (b) the generic tile kernel
    (tuning knob ntt31_off = 1) bit for bit at 2^20 x 4, in place and out of place, (c) inverse(forward(x)) == x at the domain size."""
    import os
    import port
    import torch
    fp = utils.field_params(name)
    p = fp["p"]
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, utils.to_limbs([omega(name, dom_log)], 1)[0])
    rng = random.Random(5)
    g = 0x7654321 % p
    for logn in (10, 11, 12, 13, 14, 15, 16, 17, 18, 19):
        n = 1 << logn
        batch = 3 if logn <= 12 else 1
        x = np.array([rng.randrange(p) for _ in range(n * batch)], dtype=np.uint32).reshape(-1, 1)
        w = omega(name, logn)
        for inverse, coset in ((False, 1), (True, 1), (False, g), (True, g)):
            if logn >= 17 and coset != 1 and inverse:
                continue  # keep the CPU oracle time bounded
            exp = []
            for b in range(batch):
                exp += port.ntt([int(v) for v in x[b * n:(b + 1) * n, 0]], w, p, inverse=inverse, coset=coset, field_name=name)
            cfg = ib.NTTConfig(batch_size=batch, coset_gen=utils.to_limbs([coset], 1)[0] if coset != 1 else None)
            got = ib.ntt(field, x, n, ib.NTTDir.kInverse if inverse else ib.NTTDir.kForward, cfg)
            assert got.reshape(-1).tolist() == exp, (name, logn, inverse, coset)
    # bit-reversed output (kNR: the in-place schedule, k_ntt31_inplace) is the same data permuted; inverse + coset too
    for logn, batch in ((10, 2), (13, 1), (14, 3), (16, 1), (19, 1)):
        n = 1 << logn
        x = np.array([rng.randrange(p) for _ in range(n * batch)], dtype=np.uint32).reshape(-1, 1)
        perm = np.array([common.bitrev(i, logn) for i in range(n)])
        perm_b = np.concatenate([perm + b * n for b in range(batch)])
        for d, coset in ((ib.NTTDir.kForward, 1), (ib.NTTDir.kInverse, g), (ib.NTTDir.kForward, g)):
            cg = utils.to_limbs([coset], 1)[0] if coset != 1 else None
            nn = ib.ntt(field, x, n, d, ib.NTTConfig(batch_size=batch, coset_gen=cg, ordering=ib.Ordering.kNN))
            nr = ib.ntt(field, x, n, d, ib.NTTConfig(batch_size=batch, coset_gen=cg, ordering=ib.Ordering.kNR))
            assert np.array_equal(nr, nn[perm_b]), (name, logn, d, coset)
    # (b) new pass vs the generic tile kernel
    logn, batch = 20, 4
    n = 1 << logn
    xd = torch.randint(0, p, (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32).contiguous()
    for d in (ib.NTTDir.kForward, ib.NTTDir.kInverse):
        cfg = lambda: ib.NTTConfig(batch_size=batch, is_async=False, coset_gen=utils.to_limbs([g], 1)[0])
        y_new = ib.device_empty(n * batch)
        ib.ntt(field, xd, n, d, cfg(), y_new)
        inplace = xd.clone()
        ib.ntt(field, inplace, n, d, cfg(), inplace)
        ib.set_tuning("ntt31_off", 1)
        try:
            y_old = ib.device_empty(n * batch)
            ib.ntt(field, xd, n, d, cfg(), y_old)
        finally:
            ib.set_tuning("ntt31_off", None)
        assert torch.equal(y_new.view(-1), y_old.view(-1))
        assert torch.equal(inplace.view(torch.int32).view(-1), y_new.view(torch.int32).view(-1))
    # (c) round trip at the full domain size
    n = 1 << dom_log
    xd = torch.randint(0, p, (n,), dtype=torch.int64, device="cuda").to(torch.int32).contiguous()
    y = ib.device_empty(n)
    ib.ntt(field, xd, n, ib.NTTDir.kForward, ib.NTTConfig(), y)
    z = ib.device_empty(n)
    ib.ntt(field, y, n, ib.NTTDir.kInverse, ib.NTTConfig(), z)
    assert torch.equal(z.view(torch.int32).view(-1), xd.view(-1))
    ib.ntt_release_domain(field)


@pytest.mark.parametrize("field,name", [(ib.Field.BABYBEAR, "babybear"), (ib.Field.KOALABEAR, "koalabear")])
def test_small_field_columns_batch_transposed_path(field, name):
    """columns_batch on the 4-byte fields at n >= 2^10 goes transpose -> row-batched tile pass -> transpose
    (ntt_columns_transposed): bit-identical to (a) the row-batched transform of the transposed data (itself pinned to the
    reference goldens) and (b) the strided register-only schedule it replaces (B200_NTT_COLUMNS_STRIDED=1), for ragged column
    counts, every ordering, inverse and coset, host and device buffers, in place."""
    import os
    fp = utils.field_params(name)
    p = fp["p"]
    dom_log = 14
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, utils.to_limbs([omega(name, dom_log)], 1)[0])
    rs = np.random.RandomState(31)
    g = utils.to_limbs([0x3456789 % p], 1)[0]
    for logn, cols in ((10, 2), (11, 5), (12, 33), (14, 64)):
        n = 1 << logn
        x = rs.randint(0, p, size=(n, cols), dtype=np.int64).astype(np.uint32)        # [n][cols]: column c is one transform
        xt = np.ascontiguousarray(x.T).reshape(-1, 1)                                  # [cols][n]
        for d in (ib.NTTDir.kForward, ib.NTTDir.kInverse):
            for o in (ib.Ordering.kNN, ib.Ordering.kNR, ib.Ordering.kRN, ib.Ordering.kRR):
                for cg in (None, g):
                    rows = ib.ntt(field, xt, n, d, ib.NTTConfig(batch_size=cols, ordering=o, coset_gen=cg)).reshape(cols, n)
                    got = ib.ntt(field, x.reshape(-1, 1), n, d, ib.NTTConfig(batch_size=cols, columns_batch=True, ordering=o, coset_gen=cg))
                    assert np.array_equal(got.reshape(n, cols), rows.T), (name, logn, cols, d, o, cg is not None)
        ib.set_tuning("ntt_columns_strided", 1)
        try:
            old = ib.ntt(field, x.reshape(-1, 1), n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=cols, columns_batch=True, coset_gen=g))
        finally:
            ib.set_tuning("ntt_columns_strided", None)
        dx = ib.to_device(x.reshape(-1, 1))
        ib.ntt(field, dx, n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=cols, columns_batch=True, coset_gen=g, are_outputs_on_device=True), dx)
        assert np.array_equal(ib.to_host(dx).reshape(-1, 1), old), (name, logn, cols)
    ib.ntt_release_domain(field)
