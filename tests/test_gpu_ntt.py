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
