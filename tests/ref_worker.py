"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference CPU backend (oracle/_ref/<family>) in its OWN process.

Each reference build carries its own libicicle_device.so, so only one curve/field family can be loaded per process
(oracle/ref_icicle.py).  The full-size parity tests therefore ask this worker for the reference's answer:

    python tests/ref_worker.py msm  <family> <logn> <g2 0|1> <seed> <out_prefix>
        inputs : scalars = seeded uniform values below the scalar modulus (common.seeded_scalars), points = the reference's own
                 generator (<family>_generate_affine_points: 100 distinct points repeated, curves/projective.h:37-53)
        writes : <out_prefix>_scalars.npy, _points.npy, _expected_affine.npy (reference msm -> reference to_affine), _meta.json
    python tests/ref_worker.py ntt  <family> <logn> <batch> <dir 0|1> <ordering> <seed> <out_prefix>
        inputs : seeded uniform values below the modulus; domain = get_root_of_unity(2^logn)
        writes : <out_prefix>_root.npy, _expected.npy (+ the CPU seconds in _meta.json); the caller regenerates the input
Nothing under icicle_b200/ imports this file."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import ref_icicle
    import common
    op, family = sys.argv[1], sys.argv[2]
    r = ref_icicle.get(family)
    t = ref_icicle.TARGETS[family]
    if op == "msm":
        logn, g2, seed, prefix = int(sys.argv[3]), int(sys.argv[4]) != 0, int(sys.argv[5]), sys.argv[6]
        n = 1 << logn
        s = common.seeded_scalars(common.CURVE_FIELDS[family][0], n, seed)
        P = r.generate_affine_points(n, g2=g2)
        t0 = time.perf_counter()
        res = r.msm(s, P, n, g2=g2)
        dt = time.perf_counter() - t0
        np.save(prefix + "_scalars.npy", s)
        np.save(prefix + "_points.npy", P)
        np.save(prefix + "_expected_affine.npy", r.to_affine(res[0], g2=g2))
        json.dump({"cpu_s": dt, "cores": os.cpu_count()}, open(prefix + "_meta.json", "w"))
    elif op == "ntt":
        logn, batch, d, ordering, seed, prefix = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]), sys.argv[8])
        n = 1 << logn
        fname = {"bn254": "bn254_fr", "bls12_381": "bls12_381_fr", "bls12_377": "bls12_377_fr", "bw6_761": "bls12_377_fq"}.get(family, family)
        x = common.seeded_scalars(fname, n * batch, seed)
        root = r.get_root_of_unity(n)
        r.ntt_init_domain(root)
        t0 = time.perf_counter()
        y = r.ntt(x, n, d, batch_size=batch, ordering=ordering)
        dt = time.perf_counter() - t0
        r.ntt_release_domain()
        np.save(prefix + "_root.npy", root)
        np.save(prefix + "_expected.npy", y)
        json.dump({"cpu_s": dt, "cores": os.cpu_count()}, open(prefix + "_meta.json", "w"))
    else:
        raise SystemExit("unknown op " + op)
    assert t is not None


if __name__ == "__main__":
    main()
