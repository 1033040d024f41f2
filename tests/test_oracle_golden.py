"""Pins the CPU oracle (oracle/oracle.c) against golden vectors produced by the UNMODIFIED reference
library (tests/golden/make_golden.py).  Integer stages must be bit-exact; floating-point stages are
held to the tolerances of SURVEY.md section 8(c), written next to each assertion."""
import os

import numpy as np
import pytest

from oracle import oracle as O

TOL_EXTPROD_RAW = 1 << 29      # one external product, raw coefficients (ref-vs-exact max 2^27.7)
TOL_PHASE = 1 << 44            # blind rotation, phase under the secret key


def nat_bsk(g):
    return O.permute_from_host(g["bsk_host"], g["layout"])


def test_decompose_bit_exact(golden):
    P = golden["P"]
    for j in range(P["l"]):
        got = O.decompose_i(golden["poly"], P["Bg_bit"], P["l"], j)
        assert np.array_equal(got, golden["poly_decomp"][j])


def test_rotations_bit_exact(golden):
    for a, (r1, r2) in zip(golden["rot_amounts"], golden["rot_out"]):
        assert np.array_equal(O.mul_by_xai(golden["poly"], a), r1)
        if a != 0:
            assert np.array_equal(O.mul_by_xai_minus_1(golden["poly"], a), r2)
        else:
            assert not O.mul_by_xai_minus_1(golden["poly"], a).any()


def test_slot_order_monomial(golden):
    """slot h of the host build holds p(w^e_h): transform of X is w^e_h itself."""
    N = golden["P"]["N"]
    e = O.slot_exponents(golden["layout"], N)
    h = golden["monomial_dft_host"]
    ang = np.pi * e / N
    assert np.allclose(h[: N // 2], np.cos(ang), atol=1e-12)
    assert np.allclose(h[N // 2:], np.sin(ang), atol=1e-12)


def test_slot_order_ffnt(golden_ffnt):
    g = golden_ffnt
    N = g["P"]["N"]
    assert g["layout"] == 2
    e = O.slot_exponents(2, N)
    ang = np.pi * e / N
    assert np.allclose(g["monomial_dft_host"][: N // 2], np.cos(ang), atol=1e-12)
    assert np.allclose(g["monomial_dft_host"][N // 2:], np.sin(ang), atol=1e-12)


@pytest.mark.parametrize("which", ["spq", "ffnt"])
def test_forward_transform(golden, golden_ffnt, which):
    g = golden if which == "spq" else golden_ffnt
    want = O.permute_from_host(g["digit_dft_host"], g["layout"])
    got = O.torus_to_dft(g["poly_decomp"][0].view(np.uint64))
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-12 * scale      # f64 transform of 9-bit digits


@pytest.mark.parametrize("which", ["spq", "ffnt"])
def test_external_product(golden, golden_ffnt, which):
    g = golden if which == "spq" else golden_ffnt
    P = g["P"]
    trgsw = O.permute_from_host(g["bsk_host"][0], g["layout"])
    dft = O.trgsw_mul_trlwe_dft(g["ep_in"], trgsw, P["l"], P["Bg_bit"])
    want = O.permute_from_host(g["ep_dft_host"], g["layout"])
    assert np.abs(dft - want).max() <= 2.0**-40 * np.abs(want).max()
    out = O.trlwe_from_dft(dft)
    assert np.abs(O.signed_diff(out, g["ep_out"])).max() <= TOL_EXTPROD_RAW
    # inverse transform alone, fed with the reference's own Fourier-domain result
    out2 = O.trlwe_from_dft(want)
    assert np.abs(O.signed_diff(out2, g["ep_out"])).max() <= TOL_EXTPROD_RAW
    # exact-integer oracle agrees with both
    exact = O.trgsw_mul_trlwe_exact(g["ep_in"], g["bsk_torus"][0], P["l"], P["Bg_bit"])
    assert np.abs(O.signed_diff(exact, g["ep_out"])).max() <= TOL_EXTPROD_RAW
    assert np.abs(O.signed_diff(exact, out)).max() <= TOL_EXTPROD_RAW


def test_blind_rotate_phase(golden):
    g, P = golden, golden["P"]
    bsk = nat_bsk(g)
    for b in range(g["tlwe_in"].shape[0]):
        got = O.blind_rotate(g["tv"], g["tlwe_in"][b][: P["n"]], bsk, P["l"], P["Bg_bit"])
        ph_got = O.trlwe_phase(got, g["rlwe_key"])
        ph_ref = O.trlwe_phase(g["blind_rotate_out"][b], g["rlwe_key"])
        assert np.abs(O.signed_diff(ph_got, ph_ref)).max() <= TOL_PHASE
        exact = O.blind_rotate_exact(g["tv"], g["tlwe_in"][b][: P["n"]], g["bsk_torus"], P["l"], P["Bg_bit"])
        ph_ex = O.trlwe_phase(exact, g["rlwe_key"])
        assert np.abs(O.signed_diff(ph_ex, ph_ref)).max() <= TOL_PHASE


def test_functional_bootstrap(golden):
    g, P = golden, golden["P"]
    bsk = nat_bsk(g)
    for b in range(g["tlwe_in"].shape[0]):
        wo = O.functional_bootstrap_wo_extract(g["tv"], g["tlwe_in"][b], bsk, P["l"], P["Bg_bit"], 4)
        d = O.signed_diff(O.trlwe_phase(wo, g["rlwe_key"]), O.trlwe_phase(g["fb_wo_extract_out"][b], g["rlwe_key"]))
        assert np.abs(d).max() <= TOL_PHASE
        out = O.functional_bootstrap(g["tv"], g["tlwe_in"][b], bsk, P["l"], P["Bg_bit"], 4)
        ph, ph_ref = O.tlwe_phase(out, g["ext_key"]), O.tlwe_phase(g["fb_out"][b], g["ext_key"])
        assert abs(int(np.int64(np.uint64((ph - ph_ref) % 2**64)))) <= TOL_PHASE
        # decrypted LUT value identical to the reference and to the expected slot (tests.c:1593-1602)
        want = int(g["lut_vals"][g["msgs"][b]])
        assert abs(int(np.int64(np.uint64((ph - want) % 2**64)))) <= (1 << 58)
        assert O.torus2int(ph, 6) == O.torus2int(ph_ref, 6)


def test_programmable_bootstrap(golden):
    g, P = golden, golden["P"]
    bsk = nat_bsk(g)
    prec, kappa, theta = (int(x) for x in g["pb_args"])
    for b in range(g["tlwe_in"].shape[0]):
        out = O.programmable_bootstrap(g["tv"], g["tlwe_in"][b], bsk, P["l"], P["Bg_bit"], prec, kappa, theta)
        ph, ph_ref = O.tlwe_phase(out, g["ext_key"]), O.tlwe_phase(g["pb_out"][b], g["ext_key"])
        assert abs(int(np.int64(np.uint64((ph - ph_ref) % 2**64)))) <= TOL_PHASE


def test_multivalue_clot21(golden):
    g, P = golden, golden["P"]
    tb, n_luts = (int(x) for x in g["mv_args"])
    out = O.multivalue_bootstrap_CLOT21(g["mv_tv"], g["mv_in"], nat_bsk(g), P["l"], P["Bg_bit"], tb, n_luts)
    for i in range(n_luts):
        ph, ph_ref = O.tlwe_phase(out[i], g["ext_key"]), O.tlwe_phase(g["mv_out"][i], g["ext_key"])
        assert abs(int(np.int64(np.uint64((ph - ph_ref) % 2**64)))) <= TOL_PHASE


def test_extract_bit_exact(golden):
    g = golden
    for b in range(g["tlwe_in"].shape[0]):
        for i, idx in enumerate(g["extract_idx"]):
            assert np.array_equal(O.extract_tlwe(g["fb_wo_extract_out"][b], int(idx)), g["extract_out"][b][i])


def test_keyswitch_bit_exact(golden):
    g, P = golden, golden["P"]
    for b in range(g["tlwe_in"].shape[0]):
        got = O.tlwe_keyswitch(g["fb_out"][b], g["ksk"], P["base_bit"])
        assert np.array_equal(got, g["ks_out"][b])


def test_multivalue_phases(golden_mv):
    """multivalue_bootstrap_phase1 / phase2 (bootstrap.c:232-265): phase 2 is integer-only, so fed with the
    reference's own phase-1 output it must be bit-exact; phase 1 is compared in phase."""
    g, P = golden_mv, golden_mv["P"]
    nat = O.permute_from_host(g["bsk_host"], g["layout"])
    tb, log_tb = 4, 2
    for m in range(g["mv_in"].shape[0]):
        for li, lut in enumerate(g["mv_luts"]):
            got = O.multivalue_phase2(lut, g["mv_phase1"][m], tb, log_tb)
            assert np.array_equal(got, g["mv_phase2"][m][li])
        p1 = O.multivalue_phase1(g["mv_in"][m], nat, P["N"], P["k"], P["l"], P["Bg_bit"], tb)
        for i in range(tb + 1):
            e_got, e_ref = O.extract_tlwe(p1[i], 0), O.extract_tlwe(g["mv_phase1"][m][i], 0)
            d = (O.tlwe_phase(e_got, g["ext_key"]) - O.tlwe_phase(e_ref, g["ext_key"])) % 2**64
            assert abs(int(np.int64(np.uint64(d)))) <= TOL_PHASE
        # end to end: LUT value of the encrypted message (tests.c:1816-1819)
        out = O.multivalue_phase2(g["mv_luts"][0], p1, tb, log_tb)
        want = (int(g["mv_luts"][0][m]) << 61) % 2**64
        d = (O.tlwe_phase(out, g["ext_key"]) - want) % 2**64
        assert abs(int(np.int64(np.uint64(d)))) <= (1 << 58)


def test_trlwe_table_keyswitches_bit_exact(golden_cb):
    """trlwe_priv_keyswitch (keyswitch.c:639-657) and trlwe_packing1_keyswitch (keyswitch.c:458-476): integer
    only, so bit-exact against the reference's outputs on the same random TLWE inputs."""
    g, P = golden_cb, golden_cb["P"]
    for i in range(g["ks_in"].shape[0]):
        assert np.array_equal(O.table_keyswitch_trlwe(g["ks_in"][i], g["kska"], 1, P["base_bit"]), g["priv_out"][i])
        assert np.array_equal(O.table_keyswitch_trlwe(g["ks_in"][i], g["kskb"], 0, P["base_bit"]), g["pack_out"][i])


def test_circuit_bootstrap_2(golden_cb):
    """circuit_bootstrap_2 (bootstrap.c:324-345): every TRGSW row compared in phase under the TRLWE key.  The
    blind rotation differs from the reference's by floating-point rounding, which moves a few key-switch digits
    (weight 2^(64 - t*base_bit) = 2^52 each), hence the 2^57 bound."""
    g, P = golden_cb, golden_cb["P"]
    nat = O.permute_from_host(g["bsk_host"], g["layout"])
    worst = []
    for c in range(g["cb_in"].shape[0]):
        got = O.circuit_bootstrap_2(g["cb_in"][c], nat, g["kska"], g["kskb"], P["l"], P["Bg_bit"], P["Bg_bit"], P["base_bit"])
        assert got.shape == g["cb_out"][c].shape
        for r in range(2 * P["l"]):
            d = O.signed_diff(O.trlwe_phase(got[r], g["rlwe_key"]), O.trlwe_phase(g["cb_out"][c][r], g["rlwe_key"]))
            assert np.abs(d).max() <= (1 << 57), (c, r, int(np.abs(d).max()))
            worst.append(int(np.abs(d).max()))
    # no digit moved in most rows: those agree to the rounding of the blind rotation itself
    assert sorted(worst)[len(worst) // 2] <= (1 << 40)


def test_trlwe_fft_keyswitches(golden_cb):
    """trlwe_keyswitch (keyswitch.c:162-193) and trlwe_priv_keyswitch_2 (keyswitch.c:52-63): raw coefficients
    against the reference.  Same tolerance class as one external product: t digits of base_bit bits times
    torus-sized key values, rounded in f64 (2^24 leaves a wide margin over the ~2^15 expected at N = 64)."""
    g = golden_cb
    t2, bb2 = (int(x) for x in g["params2"])
    nat = O.permute_from_host(g["kska2"], g["layout"])
    for i in range(g["rks_in"].shape[0]):
        got = O.trlwe_keyswitch(g["rks_in"][i], nat[0][None], bb2)
        assert np.abs(O.signed_diff(got, g["rks_out"][i])).max() <= (1 << 24)
        got = O.trlwe_priv_keyswitch_2(g["rks_in"][i], nat, bb2)
        assert np.abs(O.signed_diff(got, g["priv2_out"][i])).max() <= (1 << 24)


def test_circuit_bootstrap_1_and_3(golden_cb):
    """circuit_bootstrap (bootstrap.c:309-322) and circuit_bootstrap_3 (bootstrap.c:347-366), rows in phase."""
    g, P = golden_cb, golden_cb["P"]
    t2, bb2 = (int(x) for x in g["params2"])
    nat = O.permute_from_host(g["bsk_host"], g["layout"])
    nat2 = O.permute_from_host(g["kska2"], g["layout"])
    for c in range(g["cb_in"].shape[0]):
        got1 = O.circuit_bootstrap(g["cb_in"][c], nat, g["kska"], g["kskb"], P["l"], P["Bg_bit"], P["l"], P["Bg_bit"],
                                   P["base_bit"])
        got3 = O.circuit_bootstrap_3(g["cb_in"][c], nat, nat2, g["kskb"], P["l"], P["Bg_bit"], P["Bg_bit"], bb2,
                                     P["base_bit"])
        for got, want in ((got1, g["cb1_out"][c]), (got3, g["cb3_out"][c])):
            for r in range(2 * P["l"]):
                d = O.signed_diff(O.trlwe_phase(got[r], g["rlwe_key"]), O.trlwe_phase(want[r], g["rlwe_key"]))
                assert np.abs(d).max() <= (1 << 57), (c, r, int(np.abs(d).max()))


def _dft_rows_to_torus(rows_nat):
    """Natural-order Fourier rows [..., N] -> torus rows (polynomial_DFT_to_torus)."""
    flat = rows_nat.reshape(-1, rows_nat.shape[-1])
    return np.stack([O.dft_to_torus(r) for r in flat]).reshape(rows_nat.shape)


def test_trgsw_accumulator_bootstrap(golden_r4):
    """functional_bootstrap_trgsw_phase1 / phase2 (bootstrap.c:286-306).  Phase 1 is a blind rotation of every
    row of the trivial TRGSW(1): rows compared in phase (the Fourier-domain output is brought back to the torus
    first); phase 2 fed with the reference's own phase-1 output is one external product (<= 2^29 raw + extract)."""
    g, P = golden_r4, golden_r4["P"]
    nat = O.permute_from_host(g["bsk_host"], g["layout"])
    tol = max(TOL_PHASE, 1 << (64 - P["l"] * P["Bg_bit"] + 7))
    for c in range(g["r4_in"].shape[0]):
        rows_t, rows_d = O.functional_bootstrap_trgsw_phase1(g["r4_in"][c], nat, P["l"], P["Bg_bit"], P["l"], P["Bg_bit"], 4)
        ref_t = _dft_rows_to_torus(O.permute_from_host(g["trgsw_p1"][c], g["layout"]))
        for r in range(rows_t.shape[0]):
            d = O.signed_diff(O.trlwe_phase(rows_t[r], g["rlwe_key"]), O.trlwe_phase(ref_t[r], g["rlwe_key"]))
            assert np.abs(d).max() <= tol, (c, r)
        got = O.functional_bootstrap_trgsw_phase2(O.permute_from_host(g["trgsw_p1"][c], g["layout"]), g["tv"], P["l"], P["Bg_bit"])
        assert np.abs(O.signed_diff(got, g["trgsw_p2"][c])).max() <= (1 << 29)
        # end to end: decrypts to LUT[m] (tests.c:1764, tolerance 2^60)
        out = O.functional_bootstrap_trgsw_phase2(rows_d, g["tv"], P["l"], P["Bg_bit"])
        d = (O.tlwe_phase(out, g["ext_key"]) - int(g["lut"][g["msgs"][c]])) % 2**64
        assert abs(int(np.int64(np.uint64(d)))) <= (1 << 60)


def test_unfolded_blind_rotation(golden_r4):
    """bootstrap.c:124-148 with the key layout of bootstrap.c:23-48.  The per-group TRGSW is exact integer work:
    its Fourier image (multivalue_bootstrap_UBR_phase1, :151-172) is compared after the inverse transform
    (<= 2^16 raw: 64-bit values lose 11 bits in f64, then two transforms); the rotation itself in phase."""
    g, P = golden_r4, golden_r4["P"]
    u, n = g["unfolding"], P["n"]
    tol = max(TOL_PHASE, 1 << (64 - P["l"] * P["Bg_bit"] + 7))
    for c in range(g["r4_in"].shape[0]):
        a = g["r4_in"][c][:n]
        for grp in range(n // u):
            xai = O.unfold_group(a, g["su"], grp, u, P["l"])
            ref = _dft_rows_to_torus(O.permute_from_host(g["ubr_p1"][c][grp], g["layout"]))
            assert np.abs(O.signed_diff(xai, ref)).max() <= (1 << 16), (c, grp)
        got = O.blind_rotate_unfolded(g["bru_in"][c], a, g["su"], n, u, P["l"], P["Bg_bit"])
        d = O.signed_diff(O.trlwe_phase(got, g["rlwe_key"]), O.trlwe_phase(g["bru_out"][c], g["rlwe_key"]))
        assert np.abs(d).max() <= tol, c
        got = O.functional_bootstrap_unfolded_wo_extract(g["tv"], g["r4_in"][c], g["su"], u, P["l"], P["Bg_bit"], 4)
        d = O.signed_diff(O.trlwe_phase(got, g["rlwe_key"]), O.trlwe_phase(g["fbu_out"][c], g["rlwe_key"]))
        assert np.abs(d).max() <= tol, c
        # multivalue_bootstrap_UBR_phase2 (:174-190) == the same rotation + extract
        ph = O.tlwe_phase(O.extract_tlwe(got, 0), g["ext_key"])
        ph_ref = O.tlwe_phase(g["ubr_p2"][c], g["ext_key"])
        assert abs(int(np.int64(np.uint64((ph - ph_ref) % 2**64)))) <= tol
        assert abs(int(np.int64(np.uint64((ph - int(g["lut"][g["msgs"][c]])) % 2**64)))) <= (1 << 58)


# ---- extraction family of the multi-ciphertext caller (trlwe.c:554-620), fixture recorded by make_golden.py --mvx-only ----
def test_mv_extract_family_vs_reference():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_mvx.npz"))
    for si, (k, N) in enumerate(g["shapes"]):
        tr, start = g[f"s{si}_trlwe"], g[f"s{si}_start"]
        for idx in g["idx_list"]:
            assert np.array_equal(O.extract_tlwe_acc(start, tr, int(idx), +1), g[f"s{si}_extract_addto_{int(idx)}"])
            assert np.array_equal(O.extract_tlwe_acc(start, tr, int(idx), -1), g[f"s{si}_extract_subto_{int(idx)}"])
        for amount in g["amounts"]:
            assert np.array_equal(O.mv_extract_tlwe(tr, int(amount)), g[f"s{si}_mv_{int(amount)}"])
        for scale in g["scales"]:
            assert np.array_equal(O.mv_extract_tlwe_scaling(tr, int(scale)), g[f"s{si}_scaling_{int(scale)}"])
            assert np.array_equal(O.mv_extract_tlwe_scaling(tr, int(scale), start, +1), g[f"s{si}_scaling_addto_{int(scale)}"])
            assert np.array_equal(O.mv_extract_tlwe_scaling(tr, int(scale), start, -1), g[f"s{si}_scaling_subto_{int(scale)}"])
