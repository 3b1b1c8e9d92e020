"""Multi-device use of the backend.

The reference's contract is one host thread per device calling the same implementation concurrently
(docs/docs/start/architecture/multi-device.md:32-36,76; wrappers/rust/icicle-core/src/msm/tests.rs:26-40,
icicle/src/device_api.cpp:87-102).  These tests drive (a) the in-backend orchestrator b200_msm_multi_gpu / b200_ntt_multi_gpu
(one host thread per device, batch-index and point-range sharding) and (b) plain concurrent callers, one Python thread per
device (ctypes releases the GIL), against the single-device results and the reference CPU backend.  On a box with one GPU the
"devices" are [0, 0, ...]: several host threads sharing device 0 on private streams -- the same code paths and the same
thread-safety requirements (thread-local device state, per-(field, device) NTT domains behind a mutex, a shared scratch pool)."""
import threading

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu


def _devices(k):
    have = ib.get_device_count()
    return [i % have for i in range(k)]


@pytest.fixture(scope="module")
def ref():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not present")
    return ref_icicle.get("bn254")


def test_shard_range_covers_everything():
    for total, parts in ((0, 3), (1, 4), (10, 3), (1 << 26, 8), (129, 128)):
        spans = [ib.shard_range(total, parts, i) for i in range(parts)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_msm_multi_gpu_point_range_and_batch(ref):
    C = ib.Curve.BN254_G1
    ib.set_device(0)
    # one MSM split by point range over 2 and 3 host threads / devices (ragged split), incl. the chunked host pipeline
    n = (1 << 14) + 3
    s, P = ref.generate_scalars(n), ref.generate_affine_points(n)
    P[11] = 0
    exp = ref.msm(s, P, n)
    for k in (2, 3):
        got = ib.msm_multi_gpu(C, s, P, n, device_ids=_devices(k))
        assert ref.projective_eq(got[0], exp[0]), k
    # batch >= devices: batch index partitioned, shared and per-MSM bases
    n, batch = 1 << 10, 5
    s = ref.generate_scalars(n * batch)
    Ps = ref.generate_affine_points(n)
    exp = ref.msm(s, Ps, n, batch_size=batch)
    got = ib.msm_multi_gpu(C, s, Ps, n, ib.MSMConfig(batch_size=batch), device_ids=_devices(2))
    assert all(ref.projective_eq(got[b], exp[b]) for b in range(batch))
    Pn = ref.generate_affine_points(n * batch)
    exp = ref.msm(s, Pn, n, batch_size=batch, are_points_shared_in_batch=False)
    got = ib.msm_multi_gpu(C, s, Pn, n, ib.MSMConfig(batch_size=batch, are_points_shared_in_batch=False), device_ids=_devices(3))
    assert all(ref.projective_eq(got[b], exp[b]) for b in range(batch))
    # batch < devices: each MSM's point range is split over devices / batch threads, partials summed with ec_sum
    batch = 2
    exp = ref.msm(s[: n * batch], Ps, n, batch_size=batch)
    got = ib.msm_multi_gpu(C, s[: n * batch], Ps, n, ib.MSMConfig(batch_size=batch), device_ids=_devices(4))
    assert all(ref.projective_eq(got[b], exp[b]) for b in range(batch))
    # precomputed bases: the shards must use the window size the table was built for
    pf = 3
    pre = ib.msm_precompute_bases(C, Ps, n, ib.MSMConfig(precompute_factor=pf))
    exp1 = ref.msm(s[:n], Ps, n)
    got = ib.msm_multi_gpu(C, s[:n], pre, n, ib.MSMConfig(precompute_factor=pf), device_ids=_devices(2))
    assert ref.projective_eq(got[0], exp1[0])
    # device-resident inputs are refused (they already belong to one device)
    with pytest.raises(ib.IcicleError):
        ib.msm_multi_gpu(C, ib.to_device(s[:n]), Ps, n, device_ids=_devices(2))


def test_ntt_multi_gpu_batch_rows(ref):
    F = ib.Field.BN254_FR
    ib.set_device(0)
    logn, batch = 12, 7
    n = 1 << logn
    root = ref.get_root_of_unity(n)
    ref.ntt_release_domain()
    ref.ntt_init_domain(root)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, root)
    x = ref.generate_scalars(n * batch)
    for d in (0, 1):
        for o in (ib.Ordering.kNN, ib.Ordering.kNR):
            exp = ref.ntt(x, n, d, batch_size=batch, ordering=int(o))
            got = ib.ntt_multi_gpu(F, x, n, d, ib.NTTConfig(batch_size=batch, ordering=o), device_ids=_devices(3))
            assert np.array_equal(got, exp), (d, o)
    ref.ntt_release_domain()
    for dv in set(_devices(3)):
        ib.set_device(dv)
        ib.ntt_release_domain(F)
    ib.set_device(0)


def test_concurrent_host_threads_one_per_device(ref):
    """msm/tests.rs:26-40 style: every device id driven from its own host thread at the same time, MSM and NTT interleaved."""
    C, F = ib.Curve.BN254_G1, ib.Field.BN254_FR
    n = (1 << 13) + 1
    logn = 11
    root = ref.get_root_of_unity(1 << logn)
    ref.ntt_release_domain()
    ref.ntt_init_domain(root)
    work = []
    for t in range(4):
        s, P = ref.generate_scalars(n), ref.generate_affine_points(n)
        x = ref.generate_scalars(1 << logn)
        work.append((s, P, ref.msm(s, P, n), x, ref.ntt(x, 1 << logn, 0)))
    ref.ntt_release_domain()
    devs = _devices(4)
    errors = []

    def run(t):
        try:
            ib.set_device(devs[t])
            ib.ntt_init_domain(F, root)      # idempotent per (field, device), mutex-protected
            s, P, exp_msm, x, exp_ntt = work[t]
            for rep in range(3):
                got = ib.msm(C, s, P, n, ib.MSMConfig(c=(0, 9, 13)[rep]))
                if not np.array_equal(common.projective_to_affine_ints(got[0], 8, utils.field_params("bn254_fq")["p"]),
                                      common.projective_to_affine_ints(exp_msm[0], 8, utils.field_params("bn254_fq")["p"])):
                    errors.append(("msm", t, rep))
                y = ib.ntt(F, x, 1 << logn, ib.NTTDir.kForward)
                if not np.array_equal(y, exp_ntt):
                    errors.append(("ntt", t, rep))
        except Exception as e:  # noqa
            errors.append(("exc", t, repr(e)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for dv in set(devs):
        ib.set_device(dv)
        ib.ntt_release_domain(F)
    ib.set_device(0)


def _omega(fname, logn):
    fp = utils.field_params(fname)
    return pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])


@pytest.mark.parametrize("fname,field,logn", [("bn254_fr", ib.Field.BN254_FR, 16), ("babybear", ib.Field.BABYBEAR, 19), ("bls12_377_fq", ib.Field.BLS12_377_FQ, 16)])
def test_single_ntt_spanning_devices(fname, field, logn):
    """ONE transform over several devices (4-step, one all-to-all by peer copies; b200_ntt_multi_gpu with
    batch 1) must be bit-identical to the single-device transform, forward and inverse, natural order in and out."""
    L = utils.field_params(fname)["limbs"]
    n = 1 << logn
    ib.set_device(0)
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, utils.to_limbs([_omega(fname, logn)], L)[0])
    x = common.seeded_scalars(fname, n, 4242)
    for k in (2, 4):
        for d in (ib.NTTDir.kForward, ib.NTTDir.kInverse):
            exp = ib.ntt(field, x, n, d)
            got = ib.ntt_multi_gpu(field, x, n, d, device_ids=_devices(k))
            assert np.array_equal(got, exp), (fname, k, d)
    for dv in set(_devices(4)):
        ib.set_device(dv)
        ib.ntt_release_domain(field)
    ib.set_device(0)


def test_distributed_ntt_phases_with_explicit_all_to_all():
    """The building blocks the one-process-per-GPU deployment uses (bench.py: phase 1 -> NCCL all_to_all_single -> phase 2),
    driven here for G = 4 ranks inside one process with the all-to-all done by tensor slicing: column slabs of the A x B view in,
    column slabs of the B x A view of the natural-order result out; the inverse (dimensions swapped) returns the input."""
    import torch
    F, fname = ib.Field.BN254_FR, "bn254_fr"
    a_log, b_log, G = 7, 6, 4
    A, B = 1 << a_log, 1 << b_log
    n = A * B
    ib.set_device(0)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, utils.to_limbs([_omega(fname, a_log + b_log)], 8)[0])
    x = common.seeded_scalars(fname, n, 77)
    exp = ib.ntt(F, x, n, ib.NTTDir.kForward)

    def run(vec, a_log, b_log, d):
        A, B = 1 << a_log, 1 << b_log
        m = torch.from_numpy(vec.astype(np.int32)).cuda().view(A, B, 8)
        slabs = [m[:, r * (B // G):(r + 1) * (B // G), :].contiguous() for r in range(G)]
        for r in range(G):
            ib.capi.check(ib.capi.lib.b200_ntt_dist_phase1(int(F), slabs[r].data_ptr(), a_log, b_log, G, r, int(d), None), "phase1")
        torch.cuda.synchronize()
        blocks = [s.view(G, A // G, B // G, 8) for s in slabs]                       # block s of rank r = rows of rank s
        outs = []
        for s_rank in range(G):
            recv = torch.stack([blocks[r][s_rank] for r in range(G)]).contiguous()     # received in source-rank order
            out = torch.empty((B, A // G, 8), dtype=torch.int32, device="cuda")
            ib.capi.check(ib.capi.lib.b200_ntt_dist_phase2(int(F), recv.data_ptr(), out.data_ptr(), a_log, b_log, G, s_rank, int(d), None), "phase2")
            outs.append(out)
        torch.cuda.synchronize()
        full = torch.cat(outs, dim=1).contiguous()                                     # [B][A]: the natural-order result
        return full.view(-1, 8).cpu().numpy().astype(np.uint32)

    got = run(x, a_log, b_log, ib.NTTDir.kForward)
    assert np.array_equal(got, exp)
    back = run(got, b_log, a_log, ib.NTTDir.kInverse)
    assert np.array_equal(back, x)
    ib.ntt_release_domain(F)
