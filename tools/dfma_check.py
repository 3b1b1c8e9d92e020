"""Checks gpurun_out/dfma_check.bin (written by tools/dfma_modmul.cu on the GPU box) with Python integers: the FP64-pipe
Montgomery product must equal a*b*2^-260 mod p (any representative below 2p), for one level and for a squared second level."""
import struct, sys
p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
Rinv = pow(1 << 260, -1, p)
data = open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dfma_check.bin", "rb").read()
n = len(data) // (8 * 20)
vals = struct.unpack("<%dQ" % (n * 20), data)
inp, out = vals[: n * 10], vals[n * 10:]
val = lambda v: sum(x << (52 * i) for i, x in enumerate(v))
bad = 0
mx = 0
for t in range(n):
    a, b = val(inp[t * 10: t * 10 + 5]), val(inp[t * 10 + 5: t * 10 + 10])
    r, r2 = val(out[t * 10: t * 10 + 5]), val(out[t * 10 + 5: t * 10 + 10])
    ok = r < 2 * p and r % p == a * b * Rinv % p and r2 < 2 * p and r2 % p == r * r * Rinv % p
    ok = ok and all(x < (1 << 52) for x in out[t * 10: t * 10 + 10])
    mx = max(mx, r, r2)
    bad += not ok
print(f"{n} samples, {bad} mismatches, max result / p = {mx / p:.4f}")
sys.exit(1 if bad else 0)
