"""Drive tests/emul/ff_emul_test.cpp (host emulation of icicle_b200/csrc/ff.cuh) against Python integers."""
import json, os, random, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")


def run(exe, n_random=300, seed=1):
    d = json.load(open(os.path.join(ROOT, "icicle_b200", "params.json")))["fields"]
    random.seed(seed)
    lines, exp = [], []
    for name, f in d.items():
        p = int(f["p"], 16); n = f["limbs"]; R = 1 << (32 * n); Rinv = pow(R, -1, p)
        cases = [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 1), (0, p - 1), (1, 0), (p - 1, 2), ((1 << (p.bit_length() - 1)), p - 2)]
        cases += [(random.randrange(p), random.randrange(p)) for _ in range(n_random)]
        for a, b in cases:
            lines.append(f"{name} {a:x} {b:x}")
            w = 8 * n
            exp.append(" ".join(f"{x:0{w}x}" for x in [(a + b) % p, (a - b) % p, a * b * Rinv % p, a * R % p, a * Rinv % p, (-a) % p]) + " ")
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.splitlines()
    bad = [(l, o, e) for l, o, e in zip(lines, out, exp) if o != e]
    return len(lines), len(out), bad


if __name__ == "__main__":
    n, m, bad = run(sys.argv[1])
    print(n, m, "mismatches:", len(bad))
    for b in bad[:3]: print(b)
