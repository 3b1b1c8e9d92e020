"""Drive tests/emul/ec_emul_test.cpp (host emulation of icicle_b200/csrc/ec.cuh: the XYZZ group law the MSM / ECNTT kernels run)
against Python-integer curve arithmetic, including the special cases: equal points (doubling inside the add), P + (-P),
the affine zero on either side, the point at infinity as accumulator."""
import os, random, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(exe, n_random=40, seed=2):
    import common
    from icicle_b200 import utils
    random.seed(seed)
    lines, exp = [], []
    for curve, fq in (("bn254", "bn254_fq"), ("bls12_381", "bls12_381_fq"), ("grumpkin", "bn254_fr")):
        q = utils.field_params(fq)["p"]
        cp = utils.curve_params(curve)
        G = (cp["gx"], cp["gy"])
        pts = [common.ec_mul(random.randrange(1, 1 << 64), G, q) for _ in range(8)]
        cases = [(pts[0], pts[0]), (pts[1], (pts[1][0], (-pts[1][1]) % q)), (None, pts[2]), (pts[3], None), (None, None)]
        cases += [(random.choice(pts), random.choice(pts)) for _ in range(n_random)]
        for P1, P2 in cases:
            z = random.randrange(2, q)
            f = lambda P: ("0", "0") if P is None else (f"{P[0]:x}", f"{P[1]:x}")
            lines.append(" ".join((curve,) + f(P1) + f(P2) + (f"{z:x}",)))
            neg = lambda P: None if P is None else (P[0], (-P[1]) % q)
            s = common.ec_add(P1, P2, q)
            exp.append((q, [s, s, common.ec_add(P1, P1, q), common.ec_add(s, P2, q), common.ec_add(P1, neg(P1), q)]))
    n_g1 = len(lines)
    # ---- G2 over Fq2 = Fq[u]/(u^2 - nr): elements are (re, im) pairs --------------------------------------------------------
    for curve, fq in (("bn254", "bn254_fq"), ("bls12_381", "bls12_381_fq")):
        q = utils.field_params(fq)["p"]
        cp = utils.curve_params(curve)
        nr = cp["nonresidue"]
        mul = lambda a, b, q=q, nr=nr: ((a[0] * b[0] + nr * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)
        sub = lambda a, b, q=q: ((a[0] - b[0]) % q, (a[1] - b[1]) % q)
        def inv(a, q=q, nr=nr):
            d = pow((a[0] * a[0] - nr * a[1] * a[1]) % q, -1, q)
            return (a[0] * d % q, (-a[1]) * d % q)
        def add2(P1, P2, q=q, mul=mul, sub=sub, inv=inv):
            if P1 is None: return P2
            if P2 is None: return P1
            (x1, y1), (x2, y2) = P1, P2
            if x1 == x2:
                if y1 != y2 or y1 == (0, 0): return None
                xx = mul(x1, x1)
                lam = mul(((3 * xx[0]) % q, (3 * xx[1]) % q), inv(((2 * y1[0]) % q, (2 * y1[1]) % q)))
            else:
                lam = mul(sub(y2, y1), inv(sub(x2, x1)))
            x3 = sub(sub(mul(lam, lam), x1), x2)
            return (x3, sub(mul(lam, sub(x1, x3)), y1))
        def mul2(k, P, add2=add2):
            acc = None
            while k:
                if k & 1: acc = add2(acc, P)
                P = add2(P, P); k >>= 1
            return acc
        G2 = ((cp["g2_gen_x_re"], cp["g2_gen_x_im"]), (cp["g2_gen_y_re"], cp["g2_gen_y_im"]))
        pts = [mul2(random.randrange(1, 1 << 40), G2) for _ in range(5)]
        neg2 = lambda P, q=q: None if P is None else (P[0], ((-P[1][0]) % q, (-P[1][1]) % q))
        cases = [(pts[0], pts[0]), (pts[1], neg2(pts[1])), (None, pts[2]), (pts[3], None)] + [(random.choice(pts), random.choice(pts)) for _ in range(max(4, n_random // 4))]
        for P1, P2 in cases:
            z = (random.randrange(1, q), random.randrange(q))
            f = lambda P: ("0", "0", "0", "0") if P is None else tuple(f"{v:x}" for v in (P[0][0], P[0][1], P[1][0], P[1][1]))
            lines.append(" ".join((curve + "_g2",) + f(P1) + f(P2) + (f"{z[0]:x}", f"{z[1]:x}")))
            s_ = add2(P1, P2)
            exp.append(((q, mul, inv), [s_, s_, add2(P1, P1), add2(s_, P2), add2(P1, neg2(P1))]))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.splitlines()
    bad = []
    for idx, (l, o, (q_, e)) in enumerate(zip(lines, out, exp)):
        w = o.split()
        if idx < n_g1:
            if len(w) != 15:
                bad.append((l, o)); continue
            for k in range(5):
                X, Y, Z = (int(v, 16) for v in w[3 * k: 3 * k + 3])
                got = None if Z == 0 else (X * pow(Z, -1, q_) % q_, Y * pow(Z, -1, q_) % q_)
                if Z == 0 and not (X == 0 and Y != 0):
                    bad.append((l, k, "zero representative is not (0, y, 0)"))
                if got != e[k]:
                    bad.append((l, k, got, e[k]))
        else:
            _qq, mul, inv = q_
            if len(w) != 30:
                bad.append((l, o)); continue
            for k in range(5):
                v = [int(t, 16) for t in w[6 * k: 6 * k + 6]]
                X, Y, Z = (v[0], v[1]), (v[2], v[3]), (v[4], v[5])
                got = None if Z == (0, 0) else (mul(X, inv(Z)), mul(Y, inv(Z)))
                if got != e[k]:
                    bad.append((l, k, got, e[k]))
    return len(lines), len(out), bad


if __name__ == "__main__":
    n, m, bad = run(sys.argv[1])
    print(n, m, "mismatches:", len(bad))
    for b in bad[:3]: print(b)
