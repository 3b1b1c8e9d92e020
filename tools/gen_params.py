#!/usr/bin/env python3
"""Generate icicle_b200/csrc/params_gen.cuh and icicle_b200/params.json.

The moduli, roots of unity, generators and curve constants are mathematical facts of the supported fields/curves; they are
read here from the reference's parameter headers (so a typo cannot creep in) and every derived constant (Montgomery R,
R^2, -p^-1 mod 2^32, 2-adicity, b in Montgomery form ...) is computed with Python integers.  Run in the build container
(needs /root/reference); the outputs are committed.

Reference sources parsed (file under /root/reference/icicle/include/icicle):
  fields/snark_fields/{bn254,bls12_381,bls12_377,bw6_761}_{scalar,base}.h  (modulus, rou)
  fields/stark_fields/{babybear,koalabear,stark252,m31}.h                      (modulus, rou)
  curves/params/{bn254,bls12_381,bls12_377,bw6_761,grumpkin}.h             (gen_x, gen_y, weierstrass_b, is_b_neg, G2)
"""
import json, os, re, sys

REF = os.environ.get("ICICLE_REF", "/root/reference") + "/icicle/include/icicle/"
OUT_CUH = os.path.join(os.path.dirname(__file__), "..", "icicle_b200", "csrc", "params_gen.cuh")
OUT_JSON = os.path.join(os.path.dirname(__file__), "..", "icicle_b200", "params.json")


def parse_arrays(path):
    """name -> int for every `static constexpr <type> name = {0x.., ...};` holding a flat hex list."""
    txt = open(REF + path).read()
    txt = re.sub(r"//[^\n]*", "", txt)
    out = {}
    for m in re.finditer(r"static\s+constexpr\s+[\w:<>, ]+?\s+(\w+)\s*=\s*\{([0-9a-fA-Fx,\s]+)\}\s*;", txt):
        name, body = m.group(1), m.group(2)
        limbs = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", body)]
        if not limbs: continue
        val = sum(l << (32 * i) for i, l in enumerate(limbs))
        out.setdefault(name, (val, len(limbs)))
    for m in re.finditer(r"static\s+constexpr\s+bool\s+(\w+)\s*=\s*(true|false)", txt):
        out.setdefault(m.group(1), (m.group(2) == "true", 0))
    for m in re.finditer(r"static\s+constexpr\s+uint32_t\s+(\w+)\s*=\s*(\d+)", txt):
        out.setdefault(m.group(1), (int(m.group(2)), 0))
    return out


def field(name, path, ref_cfg):
    a = parse_arrays(path)
    p, n = a["modulus"]
    f = {"name": name, "p": p, "limbs": n, "src": path, "cfg": ref_cfg}
    if "rou" in a: f["rou"] = a["rou"][0]
    if "nonresidue" in a:
        f["nonresidue"] = a["nonresidue"][0]
        f["nonresidue_is_negative"] = bool(a.get("nonresidue_is_negative", (False, 0))[0])
    return f


FIELDS = [
    field("bn254_fr", "fields/snark_fields/bn254_scalar.h", "bn254::fp_config"),
    field("bn254_fq", "fields/snark_fields/bn254_base.h", "bn254::fq_config"),
    field("bls12_381_fr", "fields/snark_fields/bls12_381_scalar.h", "bls12_381::fp_config"),
    field("bls12_381_fq", "fields/snark_fields/bls12_381_base.h", "bls12_381::fq_config"),
    field("bls12_377_fr", "fields/snark_fields/bls12_377_scalar.h", "bls12_377::fp_config"),
    field("bls12_377_fq", "fields/snark_fields/bls12_377_base.h", "bls12_377::fq_config"),
    field("bw6_761_fq", "fields/snark_fields/bw6_761_base.h", "bw6_761::fq_config"),
    field("stark252", "fields/stark_fields/stark252.h", "stark252::fp_config"),
    field("babybear", "fields/stark_fields/babybear.h", "babybear::fp_config"),
    field("koalabear", "fields/stark_fields/koalabear.h", "koalabear::fp_config"),
    # Mersenne-31: vec-ops only (no NTT in the reference: icicle/cmake/features.cmake:7).  The reference's MersenneField keeps
    # values canonical and defines Montgomery form as the identity (m31.h:232-234); the kernels use the generic odd-modulus
    # Montgomery arithmetic internally and b200_convert_montgomery copies (vec_ops.cu).
    field("m31", "fields/stark_fields/m31.h", "m31::fp_config"),
]
FBY = {f["name"]: f for f in FIELDS}


def two_adicity(p):
    t, k = p - 1, 0
    while t % 2 == 0: t //= 2; k += 1
    return k


for f in FIELDS:
    p, n = f["p"], f["limbs"]
    R = 1 << (32 * n)
    f["bits"] = p.bit_length()
    f["R"] = R % p
    f["R2"] = R * R % p
    f["R3"] = R * R * R % p
    f["np0"] = (-pow(p, -1, 1 << 32)) % (1 << 32)
    f["two_adicity"] = two_adicity(p)
    if "rou" in f:
        assert pow(f["rou"], 1 << f["two_adicity"], p) == 1 and pow(f["rou"], 1 << (f["two_adicity"] - 1), p) != 1, f["name"]


def curve(name, path, fr, fq, g2_kind):
    a = parse_arrays(path)
    q = FBY[fq]["p"]
    b = a["weierstrass_b"][0] % q
    if a.get("is_b_neg", (False, 0))[0]: b = (-b) % q
    c = {"name": name, "fr": fr, "fq": fq, "src": path, "gx": a["gen_x"][0], "gy": a["gen_y"][0], "b": b, "g2": g2_kind}
    assert (c["gy"] ** 2 - c["gx"] ** 3 - b) % q == 0, name
    if g2_kind == "fq2":
        nr = FBY[fq]["nonresidue"]
        c["nonresidue"] = (-nr) % q if FBY[fq]["nonresidue_is_negative"] else nr
        for k in ("g2_gen_x_re", "g2_gen_x_im", "g2_gen_y_re", "g2_gen_y_im", "weierstrass_b_g2_re", "weierstrass_b_g2_im"):
            c[k] = a[k][0]
        for part in ("re", "im"):
            if a.get("is_b_neg_g2_" + part, (False, 0))[0]:
                c["weierstrass_b_g2_" + part] = (-c["weierstrass_b_g2_" + part]) % q
        # on-curve check in Fq[u]/(u^2 - nonresidue)
        nrv = c["nonresidue"]
        def mul(x, y): return ((x[0] * y[0] + nrv * x[1] * y[1]) % q, (x[0] * y[1] + x[1] * y[0]) % q)
        X = (c["g2_gen_x_re"], c["g2_gen_x_im"]); Y = (c["g2_gen_y_re"], c["g2_gen_y_im"])
        X3 = mul(mul(X, X), X); Y2 = mul(Y, Y)
        assert ((Y2[0] - X3[0] - c["weierstrass_b_g2_re"]) % q, (Y2[1] - X3[1] - c["weierstrass_b_g2_im"]) % q) == (0, 0), name + " g2"
    elif g2_kind == "fq":
        # bw6_761: G2 is over the same Fq (second struct in the header); parse it separately
        txt = open(REF + path).read()
        g2 = txt[txt.index("struct G2"):]
        g2 = re.sub(r"//[^\n]*", "", g2)
        def grab(nm):
            m = re.search(r"\b" + nm + r"\s*=\s*\{([0-9a-fA-Fx,\s]+)\}", g2)
            limbs = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
            return sum(l << (32 * i) for i, l in enumerate(limbs))
        c["g2_gx"], c["g2_gy"], c["g2_b"] = grab("gen_x"), grab("gen_y"), grab("weierstrass_b") % q
        if re.search(r"is_b_neg\s*=\s*true", g2): c["g2_b"] = (-c["g2_b"]) % q
        assert (c["g2_gy"] ** 2 - c["g2_gx"] ** 3 - c["g2_b"]) % q == 0, name + " g2"
    return c


CURVES = [
    curve("bn254", "curves/params/bn254.h", "bn254_fr", "bn254_fq", "fq2"),
    curve("bls12_381", "curves/params/bls12_381.h", "bls12_381_fr", "bls12_381_fq", "fq2"),
    curve("bls12_377", "curves/params/bls12_377.h", "bls12_377_fr", "bls12_377_fq", "fq2"),
    curve("bw6_761", "curves/params/bw6_761.h", "bls12_377_fq", "bw6_761_fq", "fq"),
    curve("grumpkin", "curves/params/grumpkin.h", "bn254_fq", "bn254_fr", None),
]


def limbs_of(v, n): return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
def carr(v, n): return "{" + ", ".join("0x%08xu" % x for x in limbs_of(v, n)) + "}"


def emit_cuh():
    o = []
    o.append("// GENERATED by tools/gen_params.py -- do not edit.  Field / curve constants for the sm_100a kernels.")
    o.append("// Moduli / roots of unity / generators are the public parameters of each field or curve (cross-checked against")
    o.append("// the reference headers cited per struct); everything else is derived with Python integers.")
    o.append("#pragma once\n#include <cstdint>\nnamespace b200 { namespace params {")
    for f in FIELDS:
        n, p = f["limbs"], f["p"]
        o.append(f"// {f['name']}: {f['bits']}-bit prime, {n} x u32 limbs; reference {f['cfg']} (icicle/include/icicle/{f['src']})")
        o.append(f"struct {f['name']} {{")
        o.append(f"  static constexpr int N = {n};")
        o.append(f"  static constexpr int BITS = {f['bits']};")
        o.append(f"  static constexpr int TWO_ADICITY = {f['two_adicity']};")
        o.append(f"  static constexpr uint32_t NP0 = 0x{f['np0']:08x}u;  // -p^-1 mod 2^32")
        o.append(f"  static constexpr int SPARE_BITS = {32 * n - f['bits']};")
        def acc(nm, v, cmt=""):
            o.append(f"  static __host__ __device__ constexpr uint32_t {nm}(int i) {{ constexpr uint32_t a[{n}] = {carr(v, n)}; return a[i]; }}{cmt}")
        acc("p", p, "  // modulus")
        acc("r", f["R"], "  // 2^(32N) mod p  (Montgomery one)")
        acc("r2", f["R2"], "  // R^2 mod p")
        acc("r3", f["R3"], "  // R^3 mod p")
        if "rou" in f:
            acc("rou", f["rou"], "  // primitive 2^TWO_ADICITY-th root of unity, standard form")
        o.append(f"  static constexpr bool HAS_ROU = {'true' if 'rou' in f else 'false'};")
        if "nonresidue" in f:
            o.append(f"  static constexpr uint32_t NONRESIDUE = {f['nonresidue']};  // quadratic extension: u^2 = (NONRESIDUE_IS_NEG ? -1 : 1) * NONRESIDUE")
            o.append(f"  static constexpr bool NONRESIDUE_IS_NEG = {'true' if f['nonresidue_is_negative'] else 'false'};")
        o.append("};")
    o.append("}} // namespace b200::params")
    open(OUT_CUH, "w").write("\n".join(o) + "\n")


def emit_json():
    def hx(d): return {k: (hex(v) if isinstance(v, int) and not isinstance(v, bool) and k not in ("limbs", "bits", "two_adicity", "nonresidue_is_negative") else v) for k, v in d.items()}
    json.dump({"fields": {f["name"]: hx(f) for f in FIELDS}, "curves": {c["name"]: hx(c) for c in CURVES}},
              open(OUT_JSON, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    emit_cuh(); emit_json()
    print("wrote", os.path.normpath(OUT_CUH), os.path.normpath(OUT_JSON))
