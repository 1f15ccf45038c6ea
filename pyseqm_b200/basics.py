"""Energy / force assembly with the reference's call contract (seqm/basics.py: Hamiltonian 451-533,
Energy 813-1244 ground-state branch, Force 1260-1365).  Everything between "pair list exists" and
"P, E, forces exist" is one sequence of C-ABI kernel launches on the current CUDA stream."""
import time
import warnings

import torch

from . import engine
from .Molecule import narrow_orbitals, reject_unsupported, widen_orbitals
from .seqm_functions.constants import ev_kcalpmol  # noqa: F401


def _timing(molecule, key, t0):
    if molecule.const.do_timing:
        if molecule.coordinates.is_cuda:
            torch.cuda.synchronize()
        molecule.const.timing[key].append(time.time() - t0)
        return time.time()
    return t0


def _match_orbitals(plan, V, prev_mos, e_mo):
    """Crossing matcher of basics.py:846-857: from the second forward on the same Molecule on, orbital k continues
    the previous orbital k (permutation inside the occupied / virtual blocks + sign); e_mo follows in place, e_gap
    does not (the reference computes it before the matching, basics.py:840-842)."""
    if not torch.is_tensor(prev_mos) or prev_mos.shape != V.shape or prev_mos.device != V.device:
        return V
    V, e = engine.op_mo_match(plan, V, prev_mos.detach(), e_mo[:, : plan.nmax])
    e_mo[:, : plan.nmax] = e
    return V


class Energy(torch.nn.Module):
    def __init__(self, seqm_parameters):
        super().__init__()
        reject_unsupported(seqm_parameters)
        self.seqm_parameters = seqm_parameters
        self.method = seqm_parameters["method"]
        self.Hf_flag = seqm_parameters.get("Hf_flag", True)
        self.eig = seqm_parameters.get("eig", True)
        self.eps = float(seqm_parameters["scf_eps"])
        self.sp2 = seqm_parameters.get("sp2", [False])
        self.scf_converger = seqm_parameters.get("scf_converger", [2])
        self.warm_start = bool(seqm_parameters.get("b200_eig_warm_start", True))
        self.max_iter = int(seqm_parameters.get("b200_scf_max_iter", 1000))  # reference: MAX_ITER = 1000 (scf_loop.py:29)
        self.notconverged = None

    def forward(self, molecule, learned_parameters=dict(), all_terms=False, P0=None, do_force=False, *args, **kwargs):
        plan = molecule._plan
        if learned_parameters:
            # basics.py:783-789 re-packs the parameters on every call: refresh the rows named in `learned`
            if callable(learned_parameters):
                raise NotImplementedError("callable learned_parameters need autograd through the SCF; not on the B200 path")
            plan.set_parameters({k: learned_parameters[k] for k in self.seqm_parameters.get("learned", [])})
        const = molecule.const
        if plan.large and not self.sp2[0] and not plan.eig_ok:
            raise NotImplementedError(
                f"a molecule with {plan.nmax} orbitals exceeds the eigensolver route "
                f"({plan.lib.dll.seqm_max_orbitals_eig()} orbitals): use the SP2 density, sp2=[True, eps]"
            )
        t0 = time.time()
        xyz = molecule._refresh_geometry()
        # hcore(): pair integrals + Hcore assembly
        w, hab = engine.op_pair_integrals(plan, xyz)
        H = engine.op_hcore(plan, w, hab)
        t0 = _timing(molecule, "Hcore + STO Integrals", t0)
        # density: initial guess or the caller's P0 (overwritten in place, ElectronicStructure.py:78)
        # method="PM6" without d-shell elements runs the 4-slot kernels: dense tensors are widened to 9 slots per atom
        wide = molecule.orbital_stride != 4 and not plan.d_mode
        if P0 is None:
            P = engine.op_initial_density(plan)
        else:
            P = engine.op_pack(plan, narrow_orbitals(P0, plan.molsize, molecule.orbital_stride) if wide else P0)
        # restart (P0 given, e.g. an MD step): the eigenvectors of the molecule's previous forward on this plan start
        # the first density solve; a fresh single point (P0 None) always starts cold
        C0 = molecule.__dict__.get("_C_last") if (P0 is not None and self.warm_start) else None
        if C0 is not None and C0.numel() != plan.mat_total:
            C0 = None
        if self.scf_converger[0] == 3:  # KSA (scf_forward3): scf_converger = [3, {"max_rank", "err_threshold", "T_el"}]
            if plan.d_mode or self.sp2[0] or (plan.large and not plan.eig_ok):
                raise NotImplementedError("scf_converger=[3] (KSA) runs on the eigensolver route of the sp methods (<= 256 orbitals)")
            F, Eelec, notconv, n_iter = KsaOps().scf_ksa(plan, H, w, P, self.eps, self.scf_converger[1], self.max_iter)
            Clast = None
        else:
            F, Eelec, notconv, n_iter, Clast = engine.op_scf(plan, H, w, P, self.eps, self.scf_converger, self.sp2,
                                                             warm_start=self.warm_start, want_C=True, C0=C0,
                                                             max_iter=self.max_iter)  # fmt: skip
        molecule.__dict__["_C_last"] = Clast
        molecule.n_scf_iter = n_iter
        if molecule.verbose:
            tag = {0: "scf direct step  ", 1: "scf adaptive step    ", 2: "scf pulay diis   ",
                   3: "scf KSA step     "}[self.scf_converger[0]]  # fmt: skip
            print(f"{tag}: {n_iter:>3d} | N not converged: {int(notconv.sum())}")
        if bool(notconv.any()):
            nnot = int(notconv.sum())
            print("did not converge", nnot)
            warnings.warn("SCF for %d/%d molecules doesn't converge after %d iterations" % (nnot, plan.nmol, self.max_iter))
        t0 = _timing(molecule, "SCF", t0)
        self.notconverged = notconv
        if wide or plan.d_mode:
            # (npairs, 45, 45) with the roles of the two atoms swapped (hcore.py:143-146): 16 KB per pair, so it is
            # assembled only if somebody reads molecule.w (Molecule.__getattr__)
            molecule.__dict__.pop("w", None)
            molecule.__dict__["_w_parts"] = (w, plan._wd[0] if plan.d_mode else None)
        else:
            molecule.w = w
        molecule._gam = w[:, 0, 0]
        prev_mos = molecule.molecular_orbitals  # basics.py:846: the orbitals of the previous forward on this molecule
        if self.eig and plan.large and not plan.eig_ok:
            # final eigenpairs of a molecule beyond the one-sided Jacobi kernel (e.g. C380): one cuSOLVER call per
            # molecule outside the SCF hot loop
            Fd = engine.op_unpack(plan, F)
            N = molecule.orbital_stride * plan.molsize
            e_mo = torch.zeros((plan.nmol, N), dtype=torch.float64, device=plan.device)
            V = torch.zeros((plan.nmol, plan.nmax, plan.nmax), dtype=torch.float64, device=plan.device)
            for m in range(plan.nmol):
                nh, ny = int(plan.nheavy[m]), int(plan.nhyd[m])
                idx = torch.cat([torch.arange(4 * nh, device=plan.device), 4 * nh + 4 * torch.arange(ny, device=plan.device)])
                ev, vec = torch.linalg.eigh(Fd[m][idx][:, idx])
                e_mo[m, : idx.numel()] = ev
                V[m, : idx.numel(), : idx.numel()] = vec
            lumo = plan.nocc.unsqueeze(1)
            e_gap = (e_mo.gather(1, lumo) - e_mo.gather(1, lumo - 1)).reshape(-1)
            molecule.molecular_orbitals = _match_orbitals(plan, V, prev_mos, e_mo)
        elif self.eig:
            # eigenpairs of the converged Fock matrix, warm-started from the last SCF eigenbasis
            e_mo_n, _, Cm = engine.op_eig_density(plan, F, want_P=False, want_C=True,
                                                  Cguess=Clast if self.warm_start else None)  # fmt: skip
            N = molecule.orbital_stride * plan.molsize
            e_mo = torch.zeros((plan.nmol, N), dtype=torch.float64, device=plan.device)
            e_mo[:, : plan.nmax] = e_mo_n
            lumo = plan.nocc.unsqueeze(1)
            e_gap = (e_mo.gather(1, lumo) - e_mo.gather(1, lumo - 1)).reshape(-1)
            molecule.molecular_orbitals = _match_orbitals(plan, engine.op_orbitals_dense(plan, Cm), prev_mos, e_mo)
        else:
            e_mo, e_gap = None, None
        EnucAB, Enuc = engine.op_nuclear_energy(plan, xyz, w)
        # Mulliken charges, ground-state dipole (basics.py:966-969: none for PM6 with d orbitals) and the padded force
        # tensor in one launch on the packed density
        g = engine.op_gradient(plan, xyz, P) if do_force else None
        q, dip, force = engine.op_post_scf(plan, P, xyz, g, want_dipole=not plan.d_mode)
        molecule.__dict__["_q_post"] = q
        if dip is not None:
            molecule.dipole = dip
        grad = None
        if do_force:
            grad = -force
            molecule.analytical_gradient = grad
            t0 = _timing(molecule, "Force", t0)
        Pd = engine.op_unpack(plan, P, out=P0 if (P0 is not None and P0.is_contiguous() and not wide) else None)
        if wide:
            Pd = widen_orbitals(Pd, plan.molsize, molecule.orbital_stride)
        if P0 is not None and Pd is not P0:
            P0.copy_(Pd)
            Pd = P0
        Etot = Eelec + Enuc
        Eiso, eheat = _atom_sums(plan, const)  # cached on the plan: they depend on the parameters only
        Hf = Etot - Eiso
        if self.Hf_flag:
            Hf = Hf + eheat
        self._grad, self._force = grad, force
        if all_terms:
            return Hf, Etot, Eelec, Enuc, Eiso, EnucAB, e_gap, e_mo, Pd, None, notconv
        return Eelec, EnucAB, Pd, notconv


class KsaOps:
    """Building blocks of the Krylov-subspace-approximation paths (KSA-XL-BOMD, xlbomd.py:201-341; SCF by KSA,
    scf_loop.py:1135-1381) on packed device buffers."""

    KB = 8.61739e-5  # eV/K (xlbomd.py:207)
    CANON_DM_PRT_ITER = 10  # xlbomd.py:55

    def fermi_density(self, plan, F, T_el, C0=None):
        """Fermi_Q (fermi_q.py:8-72) on packed matrices -> e (nmol, nmax), Q (packed eigenvectors), occupations f (nmol, nmax),
        mu (nmol,), D0 = 2 Q f Q^t (packed), entropy S (nmol,)."""
        beta = 1.0 / (self.KB * T_el)
        e, _, Q = engine.op_eig_density(plan, F, want_P=False, want_C=True, Cguess=C0, want_e=True)
        nocc = plan.nocc
        mask = (torch.arange(plan.nmax, device=plan.device).unsqueeze(0) < plan.norb.unsqueeze(1)).to(torch.float64)
        mu = 0.5 * (e.gather(1, nocc.unsqueeze(1) - 1) + e.gather(1, nocc.unsqueeze(1)))
        nocc_f = nocc.to(torch.float64)
        f = None
        for _ in range(64):  # Newton iteration for the chemical potential, stop test over the whole batch (fermi_q.py:47-58)
            f = torch.sigmoid(-beta * (e - mu)) * mask
            occ = f.sum(dim=1)
            docc = (beta * f * (1.0 - f)).sum(dim=1).clamp_min(1e-30)
            if bool(((nocc_f - occ).abs() <= 1e-9).all()):
                break
            mu = mu + ((nocc_f - occ) / docc).unsqueeze(1)
        D = engine.op_packed_gemm(plan, engine.op_scale_columns(plan, Q, f, 2.0), Q, tb=True)  # 2 (Q f) Q^t
        ok = (f > 1e-14) & ((1.0 - f) > 1e-14)
        p = f.masked_fill(~ok, 0.5)
        S = ((-self.KB * (p * torch.log(p) + (1.0 - p) * torch.log(1.0 - p))) * ok.to(torch.float64)).sum(dim=1)
        return e, Q, f, mu.reshape(-1).contiguous(), D, S

    def density_response(self, plan, FO1, Q, e, mu, beta, m_iter=None):
        """Canon_DM_PRT (canon_dm_prt.py:6-39) on packed matrices: first-order response of the finite-temperature density to
        the Fock perturbation FO1."""
        X = engine.op_packed_gemm(plan, Q, engine.op_packed_gemm(plan, FO1, Q), ta=True)  # Q^t FO1 Q
        engine.op_canon_prt(plan, e, mu, X, beta, m_iter or self.CANON_DM_PRT_ITER)
        return engine.op_packed_gemm(plan, engine.op_packed_gemm(plan, Q, X), Q, tb=True)  # Q X Q^t

    def krylov_kernel(self, plan, w, dDS, Q, e, mu, beta, rank, thr, m_iter=None):
        """Rank-m Krylov approximation of the kernel acting on the residual dDS = D - P (xlbomd.py:238-341, Alg. 3 of JCTC 16,
        3628): Arnoldi vectors V_k, their responses W_k = PO1(V_k) - V_k, least-squares coefficients alpha (nmol, r) with
        sum_k alpha_k W_k ~ dDS, relative residual err (nmol,).  The kernel applied to dDS is -sum_k alpha_k V_k."""
        nrm = engine.op_packed_dot(plan, dDS, dDS).sqrt()
        H0 = plan.new_mat()
        V, W = [], []
        dW = dDS
        err = torch.full((plan.nmol,), 10.0, dtype=torch.float64, device=plan.device)
        Ofull = torch.zeros((plan.nmol, rank, rank), dtype=torch.float64, device=plan.device)
        rhs_full = torch.zeros((plan.nmol, rank), dtype=torch.float64, device=plan.device)
        alpha = None
        while len(V) < rank and float(err.max()) > thr:
            v = dW.clone()
            for vj in V:  # Arnoldi orthogonalisation
                engine.op_packed_axpby(plan, -engine.op_packed_dot(plan, v, vj), vj, None, v)
            engine.op_packed_axpby(plan, None, None, engine.op_packed_dot(plan, v, v).rsqrt(), v)
            V.append(v)
            FO1 = engine.op_fock(plan, v, H0, w)  # G(dD): the Fock build without the one-electron part (G_XL_LR.py:7)
            W.append(self.density_response(plan, FO1, Q, e, mu, beta, m_iter) - v)
            dW = W[-1]
            r = len(W)
            for a in range(r):  # only the new row / column of the Gram matrix and the new right-hand side entry
                Ofull[:, a, r - 1] = Ofull[:, r - 1, a] = engine.op_packed_dot(plan, W[a], W[r - 1])
            rhs_full[:, r - 1] = engine.op_packed_dot(plan, W[r - 1], dDS)
            alpha = torch.linalg.solve(Ofull[:, :r, :r], rhs_full[:, :r].unsqueeze(-1)).squeeze(-1)
            ident = -dDS
            for a in range(r):
                engine.op_packed_axpby(plan, alpha[:, a], W[a], None, ident)
            err = engine.op_packed_dot(plan, ident, ident).sqrt() / nrm
        return alpha, V, err

    def scf_ksa(self, plan, H, w, P, eps, xl, max_iter=1000):
        """scf_forward3 (scf_loop.py:1135-1381): SCF by Krylov-subspace-approximated Newton steps on the field density at
        electronic temperature T_el, P <- P - sum_k alpha_k V_k, until |dEelec| <= eps per molecule (the only criterion).
        Every molecule is carried through every iteration (the update is masked): for a converged molecule P, hence F and its
        Fermi data, no longer change, which is what the reference's subset refresh amounts to.  CANON_DM_PRT_ITER = 8 here
        (scf_loop.py:47).  -> F, Eelec, notconverged, n_iter"""
        T_el, rank, thr = float(xl["T_el"]), int(xl["max_rank"]), float(xl["err_threshold"])
        beta = 1.0 / (self.KB * T_el)
        F = engine.op_fock(plan, P, H, w)
        Eelec = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
        err = torch.ones_like(Eelec)
        notconv = torch.ones(plan.nmol, dtype=torch.bool, device=plan.device)
        Q = None
        n_iter = 0
        while bool(notconv.any()) and n_iter < max_iter:
            n_iter += 1
            e, Q, f, mu, D, S = self.fermi_density(plan, F, T_el, Q)
            alpha, V, _ = self.krylov_kernel(plan, w, D - P, Q, e, mu, beta, rank, thr, m_iter=8)
            live = notconv.to(torch.float64)
            for a in range(len(V)):
                engine.op_packed_axpby(plan, -alpha[:, a] * live, V[a], None, P)
            F = engine.op_fock(plan, P, H, w)
            Enew = engine.op_elec_energy(plan, P, H, F)
            err = torch.where(notconv, (Enew - Eelec).abs(), err)
            Eelec = torch.where(notconv, Enew, Eelec)
            notconv = err > eps
        return F, Eelec, notconv, n_iter


class ForceXL(torch.nn.Module, KsaOps):
    """XL-BOMD energy and force for a given field density P (seqm/dynamics/xlbomd.py:73-570, non-KSA branch):
    hcore -> F(P) -> D from F (Jacobi eigensolver, or SP2 when sp2=[True, eps]) -> shadow energy
    sum D o F - 1/2 (F - h) o P -> force at fixed D and P.  No SCF.

    `forward` takes/returns the dense padded layout like the reference; `forward_packed` is the same step on
    packed device buffers (what pyseqm_b200.MolecularDynamics.XL_BOMD uses every step)."""

    def __init__(self, seqm_parameters):
        super().__init__()
        reject_unsupported(seqm_parameters)
        self.seqm_parameters = seqm_parameters
        self.Hf_flag = seqm_parameters.get("Hf_flag", True)
        self.sp2 = seqm_parameters.get("sp2", [False])
        # the warm-start eigenvectors of the previous step live on the MOLECULE, tagged with its plan (`_C_xl`): a
        # driver reused for another batch must never feed the eigensolver a guess that belongs to a different plan

    def forward_packed(self, molecule, Pp, want_e=True, learned_parameters=None, xl_bomd_params=None):
        plan = molecule._plan
        const = molecule.const
        ksa = bool(xl_bomd_params) and "max_rank" in xl_bomd_params
        if plan.d_mode:
            raise NotImplementedError("XL-BOMD with PM6 d-shell elements is not on the B200 path")
        if ksa and (self.sp2[0] or plan.large and not plan.eig_ok):
            raise NotImplementedError("KSA-XL-BOMD needs the eigenpairs of the Fock matrix: eigensolver route, <= 256 orbitals")
        if learned_parameters:  # xlbomd.py:90-116 re-packs the parameters on every step
            if callable(learned_parameters):
                raise NotImplementedError("callable learned_parameters need autograd through the SCF; not on the B200 path")
            plan.set_parameters({k: learned_parameters[k] for k in self.seqm_parameters.get("learned", [])})
        t0 = time.time()
        xyz = molecule._refresh_geometry()
        w, hab = engine.op_pair_integrals(plan, xyz)
        H = engine.op_hcore(plan, w, hab)
        t0 = _timing(molecule, "Hcore + STO Integrals", t0)
        F = engine.op_fock(plan, Pp, H, w)
        extra = {}
        if ksa:
            e_mo_n, D, extra = self._ksa_density(molecule, plan, F, Pp, w, xl_bomd_params)
        elif self.sp2[0]:
            D, _ = engine.op_sp2_density(plan, F, self.sp2[1])
            e_mo_n = None
        else:
            tag, C0 = molecule.__dict__.get("_C_xl", (None, None))
            if tag is not plan or C0 is None or C0.numel() != plan.mat_total:
                C0 = None
            e_mo_n, D, C1 = engine.op_eig_density(plan, F, want_P=True, want_C=True, Cguess=C0, want_e=want_e)
            molecule.__dict__["_C_xl"] = (plan, C1)
        t0 = _timing(molecule, "D*", t0)
        Eelec = engine.op_elec_energy_xl(plan, D, Pp, F, H)
        EnucAB, Enuc = engine.op_nuclear_energy(plan, xyz, w)
        g = engine.op_gradient_xl(plan, xyz, D, Pp)
        q, dip, force = engine.op_post_scf(plan, D, xyz, g)
        t0 = _timing(molecule, "Force", t0)
        Etot = Eelec + Enuc
        Eiso, eheat = _atom_sums(plan, const)
        Hf = Etot - Eiso + (eheat if self.Hf_flag else 0.0)
        molecule.w = w
        return dict(force=force, D=D, Hf=Hf, Etot=Etot, Eelec=Eelec, Enuc=Enuc, Eiso=Eiso, e_mo_n=e_mo_n, q=q, dipole=dip,
                    **extra)  # fmt: skip

    def _ksa_density(self, molecule, plan, F, Pp, w, xl):
        """Krylov branch of EnergyXL.forward (xlbomd.py:201-341): finite-temperature density D from the eigenpairs of F(P)
        (Fermi_Q, fermi_q.py:8-72), then the rank-m Krylov approximation of the kernel acting on D - P
        (JCTC 16, 3628 (2020), Alg. 3) with the response of every Krylov vector from canonical density-matrix perturbation
        theory (Canon_DM_PRT, canon_dm_prt.py:6-39).  Everything stays in the packed layout; the matrix work runs in
        seqm_ksa.cu / the Fock and eigensolver kernels, the (nmol, nmax) chemical-potential Newton iteration and the
        rank x rank solves are a few tiny torch calls."""
        T_el, rank, thr = float(xl["T_el"]), int(xl["max_rank"]), float(xl["err_threshold"])
        beta = 1.0 / (self.KB * T_el)
        tag, C0 = molecule.__dict__.get("_C_xl", (None, None))
        if tag is not plan or C0 is None or C0.numel() != plan.mat_total:
            C0 = None
        e, Q, f, mu1, D, S = self.fermi_density(plan, F, T_el, C0)
        molecule.__dict__["_C_xl"] = (plan, Q)
        alpha, V, err = self.krylov_kernel(plan, w, D - Pp, Q, e, mu1, beta, rank, thr)
        d2 = plan.new_mat()
        for a in range(len(V)):
            engine.op_packed_axpby(plan, -alpha[:, a], V[a], None, d2)
        return e, D, dict(EEnt=-2.0 * T_el * S, dP2dt2=d2, Krylov_Error=err, Fermi_occ=f)

    def forward(self, molecule, P, cis_amp=None, learned_parameters=dict(), xl_bomd_params=dict(), *args, **kwargs):
        if molecule.orbital_stride != 4:
            raise NotImplementedError("XL-BOMD with method='PM6' is not on the B200 path; use 'PM6_SP' for sp-only elements")
        plan = molecule._plan
        r = self.forward_packed(molecule, engine.op_pack(plan, P), learned_parameters=learned_parameters,
                                xl_bomd_params=xl_bomd_params)
        Dd = engine.op_unpack(plan, r["D"])
        molecule.dipole = r["dipole"]
        molecule.__dict__["_q_post"] = r["q"]
        N = 4 * plan.molsize
        if r["e_mo_n"] is not None:
            e = torch.zeros((plan.nmol, N), dtype=torch.float64, device=plan.device)
            e[:, : plan.nmax] = r["e_mo_n"]
            lumo = plan.nocc.unsqueeze(1)
            e_gap = (e.gather(1, lumo) - e.gather(1, lumo - 1)).reshape(-1)
        else:
            e, e_gap = None, torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
        if "dP2dt2" in r:  # KSA: entropy term, kernel-propagated second derivative, Krylov residual, occupations
            return (r["force"], Dd, r["Hf"], r["Etot"], r["Eelec"], r["Enuc"], r["Eiso"], e, e_gap, r["EEnt"],
                    engine.op_unpack(plan, r["dP2dt2"]), r["Krylov_Error"], r["Fermi_occ"])  # fmt: skip
        EEnt = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
        return (r["force"], Dd, r["Hf"], r["Etot"], r["Eelec"], r["Enuc"], r["Eiso"], e, e_gap, EEnt, None, None, None)


def _atom_sums(plan, const):
    """Per-molecule sums of the isolated-atom electronic energies (energy.py:8-23) and heats of formation."""
    cached = plan.__dict__.get("_atom_sums")
    if cached is not None and cached[0] == plan.par_version:
        return cached[1], cached[2]
    Z = plan.Z
    Eiso_atom = (
        plan.parameter("U_ss") * const.ussc[Z] + plan.parameter("U_pp") * const.uppc[Z]
        + plan.parameter("g_ss") * const.gssc[Z] + plan.parameter("g_pp") * const.gppc[Z]
        + plan.parameter("g_sp") * const.gspc[Z] + plan.parameter("g_p2") * const.gp2c[Z]
        + plan.parameter("h_sp") * const.hspc[Z]
    )  # fmt: skip
    z = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
    out = (z.clone().index_add_(0, plan.atom_mol, Eiso_atom), z.clone().index_add_(0, plan.atom_mol, const.eheat[Z]))
    plan.__dict__["_atom_sums"] = (plan.par_version, out[0], out[1])
    return out


class Force(torch.nn.Module):
    """Force.forward (basics.py:1260-1365): all three force modes of the reference (autograd,
    analytical, semi-numerical) compute the same Hellmann-Feynman gradient; one kernel serves them."""

    def __init__(self, seqm_parameters):
        super().__init__()
        self.energy = Energy(seqm_parameters)
        self.seqm_parameters = seqm_parameters

    def forward(self, molecule, learned_parameters=dict(), P0=None, do_force=True, *args, **kwargs):
        Hf, Etot, Eelec, Enuc, Eiso, _, e_gap, e, D, charge, notconverged = self.energy(
            molecule, learned_parameters=learned_parameters, all_terms=True, P0=P0, do_force=do_force
        )
        force = self.energy._force if do_force else torch.tensor([])
        return force, D, Hf, Etot, Eelec, Enuc, Eiso, e, e_gap, charge, notconverged
