"""Drive tests/emul/ff_emul_test.cpp (host emulation of icicle_b200/csrc/ff.cuh) against Python integers."""
import json, os, random, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")


def run(exe, n_random=300, seed=1):
    d = json.load(open(os.path.join(ROOT, "icicle_b200", "params.json")))["fields"]
    random.seed(seed)
    lines, exp = [], []
    for name, f in d.items():
        p = int(f["p"], 16); n = f["limbs"]; R = 1 << (32 * n); Rinv = pow(R, -1, p)
        if name == "goldilocks":  # no internal Montgomery domain (csrc/goldilocks.cuh): products are plain, to/from_mont the identity
            R = Rinv = 1
        cases = [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 1), (0, p - 1), (1, 0), (p - 1, 2), ((1 << (p.bit_length() - 1)), p - 2)]
        cases += [(random.randrange(p), random.randrange(p)) for _ in range(n_random)]
        for a, b in cases:
            lines.append(f"{name} {a:x} {b:x}")
            w = 8 * n
            exp.append(" ".join(f"{x:0{w}x}" for x in [(a + b) % p, (a - b) % p, a * b * Rinv % p, a * R % p, a * Rinv % p, (-a) % p]) + " ")
    # quartic extensions F_p[x]/(x^4 - nr) (csrc/ext4.cuh): element = 4 coefficients, low coefficient in the low 32 bits
    for base in ("babybear", "koalabear"):
        f = d[base]
        p = int(f["p"], 16); nr = int(f["nonresidue"], 16); R = 1 << 32; Rinv = pow(R, -1, p)

        def emul(x, y):
            c = [0] * 7
            for i in range(4):
                for j in range(4):
                    c[i + j] += x[i] * y[j]
            return [(c[k] + nr * (c[k + 4] if k + 4 < 7 else 0)) % p for k in range(4)]

        def einv(x):
            if not any(x):
                return [0, 0, 0, 0]
            # solve x * y = 1 by Gaussian elimination over F_p on the 4x4 multiplication matrix
            M = [[0] * 4 for _ in range(4)]
            for j in range(4):
                e = [0] * 4; e[j] = 1
                col = emul(x, e)
                for i in range(4):
                    M[i][j] = col[i]
            rhs = [1, 0, 0, 0]
            for c0 in range(4):
                piv = next(r for r in range(c0, 4) if M[r][c0] % p)
                M[c0], M[piv] = M[piv], M[c0]; rhs[c0], rhs[piv] = rhs[piv], rhs[c0]
                iv = pow(M[c0][c0], -1, p)
                M[c0] = [v * iv % p for v in M[c0]]; rhs[c0] = rhs[c0] * iv % p
                for r in range(4):
                    if r != c0 and M[r][c0]:
                        k = M[r][c0]
                        M[r] = [(a - k * b) % p for a, b in zip(M[r], M[c0])]; rhs[r] = (rhs[r] - k * rhs[c0]) % p
            return rhs

        pack = lambda x: sum(v << (32 * i) for i, v in enumerate(x))
        cases = [([0] * 4, [0] * 4), ([1, 0, 0, 0], [p - 1] * 4), ([0, 1, 0, 0], [0, 0, 0, 1]), ([5, 0, 0, 0], [7, 0, 0, 0])]
        cases += [([random.randrange(p) for _ in range(4)], [random.randrange(p) for _ in range(4)]) for _ in range(n_random // 3)]
        for a, b in cases:
            lines.append(f"ext4_{base} {pack(a):x} {pack(b):x}")
            vals = [[(x + y) % p for x, y in zip(a, b)], [(x - y) % p for x, y in zip(a, b)], [v * Rinv % p for v in emul(a, b)],
                    [v * R % p for v in a], [v * Rinv % p for v in a], einv(a)]
            exp.append(" ".join(f"{pack(v):032x}" for v in vals) + " ")
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.splitlines()
    bad = [(l, o, e) for l, o, e in zip(lines, out, exp) if o != e]
    return len(lines), len(out), bad


if __name__ == "__main__":
    n, m, bad = run(sys.argv[1])
    print(n, m, "mismatches:", len(bad))
    for b in bad[:3]: print(b)
