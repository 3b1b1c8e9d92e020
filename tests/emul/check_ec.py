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
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.splitlines()
    bad = []
    for l, o, (q, e) in zip(lines, out, exp):
        w = o.split()
        if len(w) != 15:
            bad.append((l, o)); continue
        for k in range(5):
            X, Y, Z = (int(v, 16) for v in w[3 * k: 3 * k + 3])
            got = None if Z == 0 else (X * pow(Z, -1, q) % q, Y * pow(Z, -1, q) % q)
            if Z == 0 and not (X == 0 and Y != 0):
                bad.append((l, k, "zero representative is not (0, y, 0)"))
            if got != e[k]:
                bad.append((l, k, got, e[k]))
    return len(lines), len(out), bad


if __name__ == "__main__":
    n, m, bad = run(sys.argv[1])
    print(n, m, "mismatches:", len(bad))
    for b in bad[:3]: print(b)
