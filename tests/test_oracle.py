"""CPU suite: pins the oracle.  (1) the C restatement (oracle/port) against the committed golden vectors that were generated
from the UNMODIFIED reference CPU backend (tools/make_golden.py), (2) when oracle/_ref is present, the port against the
reference itself on fresh seeded inputs, (3) the reference against first-principles integer arithmetic."""
import os
import random

import numpy as np
import pytest

from icicle_b200 import utils
import common
import port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    path = os.path.join(GOLD, f"{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"no golden fixture for {name}")
    return np.load(path)


def aff(arr, limbs):
    pts = common.affine_limbs_to_ints(arr, limbs)
    return pts[0]


@pytest.mark.parametrize("name", ["bn254", "bls12_377", "bw6_761", "grumpkin"])
def test_port_msm_vs_golden(name):
    g = _gold(name)
    cp = utils.curve_params(name)
    L = utils.field_params(cp["fq"])["limbs"]
    sl = utils.field_params(cp["fr"])["limbs"]
    sc = utils.from_limbs(g["msm_scalars"])
    pts = common.affine_limbs_to_ints(g["msm_points"], L)
    assert pts[7] is None
    for c in (4, 9):
        assert port.msm(name, sc, pts, c=c) == aff(g["msm_result_affine"], L)
    for bits in (1, 17, 100):
        assert port.msm(name, sc, pts, c=6, bitsize=bits) == aff(g[f"msm_bitsize{bits}_affine"], L)
    for b in range(3):
        assert port.msm(name, sc[b * 32:(b + 1) * 32], pts[:32], c=5) == aff(g["msm_batch3_affine"][b], L)
    # Montgomery conversion = multiplication by R = 2^(32*limbs)  (fields/params_gen.h:35-50)
    r = utils.field_params(cp["fr"])["p"]
    assert utils.from_limbs(g["scalars_montgomery"]) == [s * (1 << (32 * sl)) % r for s in sc]


@pytest.mark.parametrize("name,field", [("bn254", "bn254_fr"), ("bls12_377", "bls12_377_fr"), ("bw6_761", "bls12_377_fq"), ("stark252", "stark252")])
def test_port_ntt_and_vec_vs_golden(name, field):
    g = _gold(name)
    fp = utils.field_params(field)
    p = fp["p"]
    logn = 6
    n = 1 << logn
    root = utils.from_limbs(g["ntt_root"].reshape(1, -1))[0]
    w = pow(root, 4, p)  # domain is 2^(logn+2)
    x = utils.from_limbs(g["ntt_input"])
    perm = [common.bitrev(i, logn) for i in range(n)]
    for d in (0, 1):
        nat = port.ntt(x[:n], w, p, inverse=bool(d))
        assert nat == utils.from_limbs(g[f"ntt_d{d}_o0"])                                 # kNN
        assert [nat[perm[i]] for i in range(n)] == utils.from_limbs(g[f"ntt_d{d}_o1"])    # kNR
        xr = [x[perm[i]] for i in range(n)]                                               # kRN / kRR read bit-reversed input
        natr = port.ntt(xr, w, p, inverse=bool(d))
        assert natr == utils.from_limbs(g[f"ntt_d{d}_o2"])
        assert [natr[perm[i]] for i in range(n)] == utils.from_limbs(g[f"ntt_d{d}_o3"])
        for kind in ("dom", "arb"):
            gc = utils.from_limbs(g[f"coset_{kind}"].reshape(1, -1))[0]
            assert port.ntt(x[:n], w, p, inverse=bool(d), coset=gc) == utils.from_limbs(g[f"ntt_d{d}_coset_{kind}"])
        cols = utils.from_limbs(g[f"ntt_d{d}_batch2_cols"])
        for b in range(2):
            assert port.ntt(x[b::2][:n], w, p, inverse=bool(d)) == cols[b::2]
    a, b = utils.from_limbs(g["vec_a"]), utils.from_limbs(g["vec_b"])
    assert port.field_op(field, "add", a, b) == utils.from_limbs(g["vec_add"])
    assert port.field_op(field, "sub", a, b) == utils.from_limbs(g["vec_sub"])
    assert port.field_op(field, "mul", a, b) == utils.from_limbs(g["vec_mul"])


def test_port_field_arithmetic_all_fields():
    rng = random.Random(11)
    for name in utils.PARAMS["fields"]:
        if name == "goldilocks":
            continue  # special-form field with its own arithmetic upstream (goldilocks.h), not the generic Barrett path the port restates
        p = utils.field_params(name)["p"]
        a = [rng.randrange(p) for _ in range(100)] + [0, p - 1, p - 1]
        b = [rng.randrange(p) for _ in range(100)] + [0, p - 1, 1]
        assert port.field_op(name, "mul", a, b) == [x * y % p for x, y in zip(a, b)]
        assert port.field_op(name, "add", a, b) == [(x + y) % p for x, y in zip(a, b)]
        assert port.field_op(name, "sub", a, b) == [(x - y) % p for x, y in zip(a, b)]


@pytest.mark.parametrize("cname", ["bn254", "grumpkin", "bls12_381", "bls12_377", "bw6_761"])
def test_port_msm_vs_first_principles(cname):
    cp = utils.curve_params(cname)
    q = utils.field_params(cp["fq"])["p"]
    r = utils.field_params(cp["fr"])["p"]
    rng = random.Random(5)
    pts = common.gen_g1_points(cname, 24, 3, as_ints=True)
    pts[5] = None
    sc = [rng.randrange(r) for _ in range(24)]
    sc[0], sc[1] = 0, r - 1
    assert port.msm(cname, sc, pts, c=7) == common.msm_naive_ints(sc, pts, q)
    assert port.msm(cname, sc, pts, c=4, bitsize=13) == common.msm_naive_ints([s & 0x1FFF for s in sc], pts, q)


def test_port_vs_reference_backend_fresh_inputs():
    """Same seeded inputs through the real reference (oracle/_ref) and the port."""
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built in this environment")
    r = ref_icicle.get("bn254")
    n = 200
    s, P = r.generate_scalars(n), r.generate_affine_points(n)
    exp = common.affine_limbs_to_ints(r.to_affine(r.msm(s, P, n)[0]), 8)[0]
    assert port.msm("bn254", utils.from_limbs(s), common.affine_limbs_to_ints(P, 8), c=8) == exp
    fp = utils.field_params("bn254_fr")
    logn = 7
    root = r.get_root_of_unity(1 << logn)
    r.ntt_release_domain()
    r.ntt_init_domain(root)
    x = r.generate_scalars(1 << logn)
    w = utils.from_limbs(root.reshape(1, 8))[0]
    assert port.ntt(utils.from_limbs(x), w, fp["p"]) == utils.from_limbs(r.ntt(x, 1 << logn, 0))
    assert port.ntt(utils.from_limbs(x), w, fp["p"], inverse=True) == utils.from_limbs(r.ntt(x, 1 << logn, 1))
    r.ntt_release_domain()
    # reference vs the defining sum (first principles)
    xs = utils.from_limbs(x)[:16]
    w16 = pow(w, 1 << (logn - 4), fp["p"])
    assert port.ntt(xs, w16, fp["p"]) == common.ntt_naive_ints(xs, w16, fp["p"])


def test_port_vs_golden_other_curves():
    """bls12_381 (G1 MSM + Fr NTT) and babybear (NTT) fixtures generated from their reference builds."""
    g = _gold("bls12_381")
    sc = utils.from_limbs(g["msm_scalars"])
    pts = common.affine_limbs_to_ints(g["msm_points"], 12)
    assert port.msm("bls12_381", sc, pts, c=7) == aff(g["msm_result_affine"], 12)
    for name, field in (("bls12_381", "bls12_381_fr"), ("babybear", "babybear")):
        g = _gold(name)
        fp = utils.field_params(field)
        L = fp["limbs"]
        root = utils.from_limbs(g["ntt_root"].reshape(1, -1))[0]
        w = pow(root, 4, fp["p"])
        x = utils.from_limbs(g["ntt_input"].reshape(-1, L))
        for d in (0, 1):
            assert port.ntt(x[:64], w, fp["p"], inverse=bool(d), field_name=field) == utils.from_limbs(g[f"ntt_d{d}_o0"].reshape(-1, L))
            gc = utils.from_limbs(g["coset_arb"].reshape(1, -1))[0]
            assert port.ntt(x[:64], w, fp["p"], inverse=bool(d), coset=gc, field_name=field) == utils.from_limbs(g[f"ntt_d{d}_coset_arb"].reshape(-1, L))


@pytest.mark.parametrize("name", ["babybear", "koalabear"])
def test_port_ntt_small_fields_vs_reference_golden(name):
    """tests/golden/<field>_ntt_big.npz (tools/make_golden_smallfield.py, outputs of the unmodified reference CPU backend):
    the C restatement must reproduce them bit for bit -- full outputs at 2^10 / 2^11, SHA-256 of the output bytes at 2^13."""
    import hashlib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}_ntt_big.npz"))
    fp = utils.field_params(name)
    p = fp["p"]
    root = int(g["ntt_root"][0])
    dom_log = int(g["dom_log"][0])
    coset = int(g["coset_arb"][0])
    for logn, batch in ((10, 1), (11, 3), (13, 2)):
        n = 1 << logn
        rs = np.random.RandomState(1000 + logn)
        x = rs.randint(0, p, size=(batch << logn, 1), dtype=np.int64).astype(np.uint32)
        w = pow(root, 1 << (dom_log - logn), p)
        for d in (0, 1):
            for c in (0, 1):
                rows = [port.ntt([int(v) for v in x[b * n:(b + 1) * n, 0]], w, p, inverse=bool(d), coset=coset if c else 1, field_name=name)
                        for b in range(batch)]
                y = np.array([v for row in rows for v in row], dtype=np.uint32).reshape(-1, 1)
                key = f"l{logn}_b{batch}_d{d}_c{c}"
                assert hashlib.sha256(y.tobytes()).digest() == g["sha_" + key].tobytes(), key
                if "out_" + key in g.files:
                    assert np.array_equal(y, g["out_" + key]), key


@pytest.mark.parametrize("name", ["babybear", "koalabear"])
def test_port_extension_ntt_vs_reference_golden(name):
    """tests/golden/<field>_ext_ntt.npz (tools/make_golden_ext.py, `<field>_extension_ntt` of the unmodified reference CPU
    backend built with EXT_FIELD): the quartic-extension NTT multiplies by base-field twiddles, i.e. it is the base-field NTT
    of the C restatement applied to each of the 4 coefficient planes -- row batches, columns batches, kNN and kNR."""
    import hashlib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}_ext_ntt.npz"))
    p = utils.field_params(name)["p"]
    root, dom_log, coset = int(g["ntt_root"][0]), int(g["dom_log"][0]), int(g["coset_arb"][0])
    for logn, batch, col, ordering in g["cases"].tolist():
        if logn > 12:
            continue
        n = 1 << logn
        rs = np.random.RandomState(4000 + logn)
        x = rs.randint(0, p, size=(batch << logn, 4), dtype=np.int64).astype(np.uint32)
        # element (i, b) of transform b: row-major batch -> x[b*n + i], columns batch -> x[i*batch + b]
        X = x.reshape(batch, n, 4) if not col else x.reshape(n, batch, 4).transpose(1, 0, 2)
        w = pow(root, 1 << (dom_log - logn), p)
        perm = np.arange(n)
        if ordering == 1 and logn > 0:   # kNR: output index bit-reversed
            perm = np.array([int(format(i, f"0{logn}b")[::-1], 2) for i in range(n)])
        for d in (0, 1):
            for c in (0, 1):
                Y = np.zeros_like(X)
                for b in range(batch):
                    for k in range(4):
                        col_out = port.ntt([int(v) for v in X[b, :, k]], w, p, inverse=bool(d), coset=coset if c else 1, field_name=name)
                        Y[b, :, k] = np.array(col_out, dtype=np.uint32)[perm]
                y = Y.reshape(-1, 4) if not col else np.ascontiguousarray(Y.transpose(1, 0, 2)).reshape(-1, 4)
                key = f"l{logn}_b{batch}_c{col}_o{ordering}_d{d}_g{c}"
                assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).digest() == g["sha_" + key].tobytes(), key


def test_m31_golden_vs_port_and_integers():
    """tests/golden/m31.npz (tools/make_golden_m31.py, the reference's MersenneField): canonical results in [0, p), Montgomery
    conversion = identity (m31.h:232-234); the C restatement's Barrett path and plain Python integers agree with it."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "m31.npz"))
    p = utils.field_params("m31")["p"]
    assert p == (1 << 31) - 1
    a, b = [int(v) for v in g["vec_a"][:, 0]], [int(v) for v in g["vec_b"][:, 0]]
    for op, f in (("vector_add", lambda x, y: (x + y) % p), ("vector_sub", lambda x, y: (x - y) % p), ("vector_mul", lambda x, y: x * y % p),
                  ("vector_accumulate", lambda x, y: (x + y) % p)):
        assert [int(v) for v in g[op][:, 0]] == [f(x, y) for x, y in zip(a, b)], op
    for op in ("add", "sub", "mul"):
        assert port.field_op("m31", op, a[:512], b[:512]) == [int(v) for v in g["vector_" + op][:512, 0]], op
    assert np.array_equal(g["to_montgomery"], g["vec_a"]) and np.array_equal(g["from_montgomery"], g["vec_a"])


def test_ecntt_golden_vs_definition():
    """tests/golden/bn254_ecntt.npz (tools/make_golden_ecntt.py, `bn254_ecntt` of the unmodified reference) against the defining
    sums with Python-integer group arithmetic: forward out[k] = sum_i w^(ik) (g^i P_i), inverse out[i] = g^-i N^-1 sum_k w^(-ik) P_k,
    kNR = bit-reversed output, kRN = bit-reversed input, columns batch = strided transforms (ntt_cpu.h:69-232 with E = projective_t)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bn254_ecntt.npz"))
    q = utils.field_params("bn254_fq")["p"]
    r = utils.field_params("bn254_fr")["p"]
    n, batch, logn = 16, 2, 4
    root = utils.from_limbs(g["ntt_root"].reshape(1, 8))[0]
    w = pow(root, 1 << (int(g["dom_log"][0]) - logn), r)
    coset = utils.from_limbs(g["coset"].reshape(1, 8))[0]
    pts = [common.projective_to_affine_ints(p_, 8, q) for p_ in g["input_projective"]]
    assert pts[3] is None                                             # the point at infinity in the input
    rev = [common.bitrev(i, logn) for i in range(n)]

    def transform(x, inverse, cg):
        if not inverse:
            xs = [common.ec_mul(pow(cg, i, r), x[i], q) if x[i] is not None else None for i in range(n)]
            return [common.msm_naive_ints([pow(w, i * k, r) for i in range(n)], xs, q) for k in range(n)]
        wi, ni, gi = pow(w, -1, r), pow(n, -1, r), pow(cg, -1, r)
        out = []
        for i in range(n):
            acc = common.msm_naive_ints([pow(wi, i * k, r) for k in range(n)], x, q)
            out.append(common.ec_mul(ni * pow(gi, i, r) % r, acc, q) if acc is not None else None)
        return out

    for d in (0, 1):
        for o in (0, 1, 2):
            for c in (0, 1):
                exp = common.affine_limbs_to_ints(g[f"d{d}_o{o}_g{c}_affine"], 8)
                for b in range(batch):
                    x = pts[b * n:(b + 1) * n]
                    if o == 2:
                        x = [x[rev[i]] for i in range(n)]            # kRN: logical element i sits at position rev(i)
                    y = transform(x, bool(d), coset if c else 1)
                    if o == 1:
                        y = [y[rev[i]] for i in range(n)]            # kNR: position i holds frequency rev(i)
                    assert y == exp[b * n:(b + 1) * n], (d, o, c, b)
    exp = common.affine_limbs_to_ints(g["d0_cols_affine"], 8)
    for b in range(batch):
        y = transform([pts[i * batch + b] for i in range(n)], False, 1)
        assert y == [exp[i * batch + b] for i in range(n)], b


def _ext4(name):
    fp = utils.field_params(name)
    p, nr = fp["p"], fp["nonresidue"]

    def mul(x, y):
        c = [0] * 8
        for i in range(4):
            for j in range(4):
                c[i + j] += x[i] * y[j]
        return [(c[k] + nr * c[k + 4]) % p for k in range(4)]
    return p, mul


@pytest.mark.parametrize("name", ["babybear", "koalabear"])
def test_extension_vec_ops_golden_vs_integers(name):
    """tests/golden/<field>_ext_ops.npz (tools/make_golden_extops.py, the reference's `<field>_extension_*` vec-ops) against
    first-principles arithmetic in F_p[x]/(x^4 - nr) on Python integers: pins the fixtures the GPU extension vec-ops are tested
    with (quartic_extension.h:182-199 product, :246-283 inverse, :78-88 coefficient-wise Montgomery form)."""
    g = _gold(f"{name}_ext_ops")
    p, mul = _ext4(name)
    a, b, s = g["a"].astype(object), g["b"].astype(object), [int(v) for v in g["s"].reshape(-1)]
    rows = lambda arr: [[int(v) for v in r] for r in arr]
    A, B = rows(a), rows(b)
    assert rows(g["vector_add"]) == [[(x + y) % p for x, y in zip(u, v)] for u, v in zip(A, B)]
    assert rows(g["vector_sub"]) == [[(x - y) % p for x, y in zip(u, v)] for u, v in zip(A, B)]
    assert rows(g["vector_accumulate"]) == rows(g["vector_add"])
    assert rows(g["vector_mul"]) == [mul(u, v) for u, v in zip(A, B)]
    assert rows(g["vector_mixed_mul"]) == [[x * k % p for x in u] for u, k in zip(A, s)]
    inv = rows(g["vector_inv"])
    for v, iv in zip(B, inv):
        assert (mul(v, iv) == [1, 0, 0, 0]) if any(v) else (iv == [0, 0, 0, 0])
    assert rows(g["vector_div"]) == [mul(u, iv) for u, iv in zip(A, inv)]
    n, batch = len(A) // 3, 3
    for tag, idx in (("rows", lambda bi, i: bi * n + i), ("cols", lambda bi, i: i * batch + bi)):
        add, sub_, mul_ = rows(g[f"scalar_add_vec_{tag}"]), rows(g[f"scalar_sub_vec_{tag}"]), rows(g[f"scalar_mul_vec_{tag}"])
        sm, pr = rows(g[f"vector_sum_{tag}"]), rows(g[f"vector_product_{tag}"])
        for bi in range(batch):
            acc_s, acc_p = [0, 0, 0, 0], [1, 0, 0, 0]
            for i in range(n):
                t = idx(bi, i)
                assert add[t] == [(x + y) % p for x, y in zip(A[bi], B[t])]
                assert sub_[t] == [(x - y) % p for x, y in zip(A[bi], B[t])]
                assert mul_[t] == mul(A[bi], B[t])
                acc_s = [(x + y) % p for x, y in zip(acc_s, A[t])]
                acc_p = mul(acc_p, A[t])
            assert sm[bi] == acc_s and pr[bi] == acc_p
    R = 1 << 32
    assert rows(g["convert_montgomery_1"]) == [[x * R % p for x in u] for u in A]
    assert rows(g["convert_montgomery_0"]) == [[x * pow(R, -1, p) % p for x in u] for u in A]
    assert rows(g["bit_reverse"]) == [A[common.bitrev(i, 5)] for i in range(32)]
    assert rows(g["matrix_transpose_6x8"]) == [A[r * 8 + c] for c in range(8) for r in range(6)]
    assert rows(g["slice_3_4_10"]) == [A[3 + 4 * i] for i in range(10)]


def test_goldilocks_golden_vs_defining_sum():
    """tests/golden/goldilocks.npz (reference build of the special-form field) against the defining DFT sums and plain integer
    arithmetic: pins the fixture the GPU Goldilocks path is tested with (the C port does not cover this field)."""
    g = _gold("goldilocks")
    fp = utils.field_params("goldilocks")
    p = fp["p"]
    n = 64
    root = utils.from_limbs(g["ntt_root"].reshape(1, -1))[0]
    w = pow(root, 4, p)
    x = utils.from_limbs(g["ntt_input"])
    assert utils.from_limbs(g["ntt_d0_o0"]) == common.ntt_naive_ints(x[:n], w, p)
    assert utils.from_limbs(g["ntt_d1_o0"]) == common.ntt_naive_ints(x[:n], w, p, inverse=True)
    gc = utils.from_limbs(g["coset_arb"].reshape(1, -1))[0]
    assert utils.from_limbs(g["ntt_d0_coset_arb"]) == common.ntt_naive_ints(x[:n], w, p, coset=gc)
    a, b = utils.from_limbs(g["vec_a"]), utils.from_limbs(g["vec_b"])
    assert utils.from_limbs(g["vec_mul"]) == [u * v % p for u, v in zip(a, b)]
    assert utils.from_limbs(g["vec_add"]) == [(u + v) % p for u, v in zip(a, b)]
