"""Pins the numpy oracle (oracle/seqm_oracle) to the reference: the .npz fixtures were produced by the
unmodified reference (tools/make_golden.py) and ref_json/*.json are the reference's own test goldens
(SURVEY 8(c)).  CPU only."""
import json
import os

import numpy as np
import pytest

import seqm_oracle as so
from conftest import GOLDEN, TOL_DM, TOL_E, TOL_F, load_golden

XYZ = os.path.join(GOLDEN, "xyz")

CASES = [f"cfg1_{m}_{c}" for m in ("AM1", "PM3", "MNDO", "PM6_SP") for c in ("c2", "c1", "c0")] + [
    "cfg2_PM6_SP_24", "pm6_sp_elements_c2", "opt_charged_AM1", "opt_learned_PM3", "opt_flags_MNDO", "opt_cutoff_AM1", "thirdrow_PM3_c2", "thirdrow_AM1_c2", "thirdrow_MNDO_c2", "thirdrow_PM6_SP_c2",
    "cfg1_AM1_sp2", "ref_batch_single_point_am1", "ref_ground_force_methanal", "cfg2_PM3_48", "cfg3_coronene_AM1",
]  # fmt: skip


# MNDO PH3: the pseudo-inverse of a cond=1e12 EMAT (scf_loop.py:1024-1033) feeds rounding noise of the eigensolver
# into F at the 1e-4 level, so the number of iterations to convergence is not reproducible between LAPACK builds
CHAOTIC_DIIS = {"thirdrow_MNDO_c2"}


@pytest.mark.parametrize("name", CASES)
def test_single_point_matches_reference(name):
    g = load_golden(name)
    learned = {k[len("learned_"):]: g[k] for k in g if k.startswith("learned_")}
    out = so.single_point(g["species"], g["coordinates"], g["seqm_parameters"], charges=g.get("charges", 0),
                          learned_parameters=learned)
    if name not in CHAOTIC_DIIS:
        assert out["n_scf_iter"] == g["n_scf_iter"]
    assert not out["notconverged"].any() and not g["notconverged"].any()
    for k in ("Etot", "Hf", "Eelec", "Enuc", "Eiso"):
        assert np.abs(out[k] - g[k]).max() < TOL_E, k
    # the third-row set runs through ill-conditioned (cond up to 1e12-1e16) DIIS solves that amplify
    # LAPACK-vs-LAPACK rounding to ~2e-7 in P (energies still agree to 1e-12): looser density check there
    tol_dm = 1e-6 if name.startswith("thirdrow") else TOL_DM
    assert np.abs(out["dm"] - g["dm"]).max() < tol_dm
    tol_orb = 1e-5 if name.startswith("thirdrow") else TOL_E  # orbital energies are first order in that noise
    if "e_gap" in g:  # absent with eig=False
        assert np.abs(out["e_gap"] - g["e_gap"]).max() < tol_orb
    if "e_mo" in g:
        assert np.abs(out["e_mo"] - g["e_mo"]).max() < tol_orb
    assert np.abs(out["q"] - g["q"]).max() < tol_dm
    assert np.abs(out["force"] - g["force"]).max() < TOL_F


def test_autograd_force_mode_of_reference():
    """The reference's default (autograd) forces equal the Hellmann-Feynman gradient the oracle takes."""
    g = load_golden("cfg1_AM1_autograd")
    out = so.single_point(g["species"], g["coordinates"], g["seqm_parameters"])
    assert np.abs(out["force"] - g["force"]).max() < TOL_F
    assert np.abs(out["Etot"] - g["Etot"]).max() < TOL_E


def test_operator_level_outputs():
    """hcore -> (M, w), fock(X), sym_eig_trunc, SP2 against the reference's operators (SURVEY 8(b) level B)."""
    from seqm_oracle.density import density_from_fock, packed_index, sp2_packed
    from seqm_oracle.hamiltonian import build_fock, build_hcore, hcore_upper

    for method in ("AM1", "PM3", "MNDO", "PM6_SP"):
        g = load_golden(f"cfg1_{method}_c2")
        P = so.parse(g["species"], g["coordinates"])
        par = so.method_parameters(method, P.Z)
        hc = build_hcore(P, par)
        assert np.abs(hc["w"] - g["op_w"]).max() < 1e-12
        assert np.abs(hcore_upper(hc["H"], P) - g["op_M"]).max() < 1e-12
        assert np.abs(hc["rho0i"] - g["op_rho0i"]).max() < 1e-13
        F = build_fock(P, par, hc["H"], hc["w"], g["op_X"])
        assert np.abs(F - g["op_F"]).max() < 1e-11
        D, E, _ = density_from_fock(g["op_F"], P.nHeavy, P.nHydro, P.nocc)
        assert np.abs(D - g["op_P"]).max() < 1e-10
        assert np.abs(E - g["op_e"]).max() < 1e-10
        nmax = g["op_sp2_packed"].shape[1]
        for m in range(P.nmol):
            idx = packed_index(int(P.nHeavy[m]), int(P.nHydro[m]))
            n = idx.shape[0]
            a = np.zeros((nmax, nmax))  # the reference's pack() zero-pads to the batch maximum (pack.py:76-77)
            a[:n, :n] = g["op_F"][m][np.ix_(idx, idx)]
            d, _ = sp2_packed(a, float(P.nocc[m]), 1.0e-5)
            assert np.abs(d - g["op_sp2_packed"][m]).max() < 1e-9


def _json(name):
    with open(os.path.join(GOLDEN, "ref_json", name + ".json")) as f:
        return json.load(f)


@pytest.mark.parametrize("method", ["MNDO", "AM1", "PM3", "PM6_SP"])
def test_reference_json_smoke_single_point(method):
    """tests/unit/test_smoke_single_point.py:11-41 of the reference, against its own JSON."""
    ref = _json(f"smoke_single_point_{method}")
    s, c = so.read_xyz([os.path.join(XYZ, "methane.xyz")])
    out = so.single_point(s, c, {"method": method, "scf_eps": 1.0e-6, "scf_converger": [1]})
    assert abs(out["Etot"][0] - ref["Etot"]) < 1e-5
    assert abs(out["Eelec"][0] - ref["Eelec"]) < 1e-5
    assert abs(out["Enuc"][0] - ref["Enuc"]) < 1e-5
    assert np.abs(out["force"] - np.asarray(ref["force"])).max() < 1e-5


def test_reference_json_batch_single_point_am1():
    """tests/unit/test_batch_single_point.py:10-30."""
    ref = _json("batch_single_point_am1")
    s, c = so.read_xyz([os.path.join(XYZ, "methane.xyz"), os.path.join(XYZ, "benzene.xyz")])
    out = so.single_point(s, c, {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [1]})
    assert np.abs(out["Etot"] - np.asarray(ref["Etot"])).max() < 1e-5
    assert np.abs(out["force"] - np.asarray(ref["force"])).max() < 1e-4


@pytest.mark.parametrize(
    "name,files",
    [
        ("ground_force_methane", ["methane.xyz"]),
        ("ground_force_batch_methanal", ["methanal.1.xyz", "methanal.2.xyz", "methanal.3.xyz"]),
        ("ground_force_batch_mixed", ["methane.xyz", "benzene.xyz"]),
    ],
)
def test_reference_json_ground_forces(name, files):
    """tests/unit/test_force_methods.py:62-95."""
    ref = _json(name)
    s, c = so.read_xyz([os.path.join(XYZ, f) for f in files])
    out = so.single_point(s, c, {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [1]})
    assert np.abs(out["force"] - np.asarray(ref["force"])).max() < 1e-5


def test_input_validation_matches_reference_errors():
    """Molecule.py:188-206 (unsorted rows) and basics.py:297-299 (odd electrons)."""
    s = np.array([[1, 6, 1, 1, 1]])
    c = np.zeros((1, 5, 3))
    with pytest.raises(ValueError, match="non-increasing"):
        so.parse(s, c)
    s = np.array([[6, 1, 1, 1, 0]])
    c = np.random.default_rng(0).normal(size=(1, 5, 3))
    with pytest.raises(ValueError, match="closed shell"):
        so.parse(s, c)


@pytest.mark.parametrize("name", ["op_momatch_mixed", "op_momatch_uniform"])
def test_mo_crossing_matcher_matches_reference(name):
    """oracle restatement of basics.py:596-719 against the reference's own outputs on crafted orbital sets (row argmax
    and greedy-repair routes, block-internal permutations, sign flips): permuted copies, so exact."""
    g = load_golden(name)
    V, e = so.match_orbitals(g["V_new"], g["V_old"], g["nocc"], g["norb"], g["e"])
    assert np.array_equal(e, g["e_out"])
    assert np.array_equal(V, g["V_out"])


def test_rotation_invariance():
    """tests/unit/test_invariants.py:28-55: Etot invariant, forces co-rotate (z rotation by 0.7 rad)."""
    s, c = so.read_xyz([os.path.join(XYZ, "methane.xyz")])
    sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [2]}
    a = so.single_point(s, c, sp)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    b = so.single_point(s, c @ R.T, sp)
    assert abs(a["Etot"][0] - b["Etot"][0]) < 1e-7
    assert np.abs(a["force"] @ R.T - b["force"]).max() < 1e-5


def test_installed_reference_reproduces_its_golden():
    """oracle/_ref (the unmodified reference installed by oracle/ref_runner.py) is what bench.py's reference arm and
    parity-at-size check run: it must reproduce the fixture the same code generated from /root/reference."""
    import ref_runner

    if not ref_runner.reference_available():
        pytest.skip("oracle/_ref not installed (build() installs it where /root/reference exists)")
    g = load_golden("cfg1_AM1_c2")
    out, _ = ref_runner.run_reference(g["species"], g["coordinates"], g["seqm_parameters"])
    assert out["n_scf_iter"] == int(g["n_scf_iter"])
    assert np.abs(out["Etot"] - g["Etot"]).max() < 1e-9
    assert np.abs(out["force"] - g["force"]).max() < 1e-8


# ---- PM6 with d orbitals: oracle/seqm_oracle/pm6d.py against tools/make_golden_pm6d.py fixtures -------------
PM6D_CASES = ["pm6d_organics_c1", "pm6d_organics_c2", "pm6d_organics_c0", "pm6d_diatomics_rotated", "pm6d_cfg5_16",
              "pm6d_notebook_diatomics"]  # fmt: skip


def _pm6d_setup(g):
    from seqm_oracle import pm6d
    from seqm_oracle.tables import method_parameters

    P = so.parse(g["species"], g["coordinates"])
    par = method_parameters("PM6", P.Z)
    mpd = pm6d.atom_multipoles_spd(P.Z, par, so.atom_multipoles(P.Z, par))
    return pm6d, P, par, mpd


@pytest.mark.parametrize("name", ["pm6d_organics_c1", "pm6d_diatomics_rotated", "pm6d_notebook_diatomics"])
def test_pm6d_operators_match_reference(name):
    """45 x 45 two-centre integrals (generic point-charge multipole engine vs the reference's unrolled formulas), spd
    overlaps, Hcore and the 9 x 9 Fock build incl. the one-centre d integrals derived from Slater-Condon factors."""
    g = load_golden(name)
    pm6d, P, par, mpd = _pm6d_setup(g)
    hc = pm6d.build_hcore_spd(P, par, mpd)
    assert np.abs(hc["w"] - g["op_w"].transpose(0, 2, 1)).max() < 1e-9  # the reference stores [j-pair, i-pair]
    assert np.abs(hc["di"] - g["op_di"]).max() < 1e-12
    m = P.molsize
    Mref = g["op_M"].reshape(P.nmol, m, m, 9, 9).transpose(0, 1, 3, 2, 4).reshape(P.nmol, 9 * m, 9 * m)
    Href = np.triu(Mref) + np.triu(Mref, 1).transpose(0, 2, 1)
    live = (np.arange(9)[None, None, :] < pm6d.norb_of(P.species)[:, :, None]).reshape(P.nmol, 9 * m)
    msk = live[:, :, None] & live[:, None, :]  # the reference leaves garbage in the phantom p/d slots of H and sp atoms
    assert np.abs((hc["H"] - Href) * msk).max() < 1e-9
    F = pm6d.build_fock_spd(P, par, hc["H"], hc["w"], g["op_X"])
    assert np.abs((F - g["op_F"]) * msk).max() < 1e-9


@pytest.mark.parametrize("name", PM6D_CASES)
def test_pm6d_single_point_matches_reference(name):
    g = load_golden(name)
    out = so.single_point(g["species"], g["coordinates"], g["seqm_parameters"])
    assert out["n_scf_iter"] == g["n_scf_iter"]
    assert not out["notconverged"].any()
    for k in ("Etot", "Hf", "Eelec", "Enuc", "Eiso", "e_gap"):
        assert np.abs(out[k] - g[k]).max() < 1e-9, k
    assert np.abs(out["q"] - g["q"]).max() < 1e-9
    if name != "pm6d_notebook_diatomics":  # Ti2 / S2 on an axis: degenerate frontier orbitals, P is not unique
        assert np.abs(out["dm"] - g["dm"]).max() < 1e-8
    assert np.abs(out["force"] - g["force"]).max() < 2e-6  # oracle: central differences; reference: autograd


def test_pm6d_reference_own_golden():
    """tests/reference/pm6_batch_notebook.json of the reference (S2, Ti2, TiS, BrCl, CrTi), its tolerances (1e-5)."""
    with open(os.path.join(GOLDEN, "ref_json", "pm6_batch_notebook.json")) as f:
        ref = json.load(f)
    g = load_golden("pm6d_notebook_diatomics")
    out = so.single_point(g["species"], g["coordinates"], g["seqm_parameters"])
    assert np.allclose(out["Etot"], ref["Etot"], rtol=1e-5, atol=1e-5)
    assert np.allclose(out["force"], np.asarray(ref["force"]), rtol=1e-5, atol=1e-5)
