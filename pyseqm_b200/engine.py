"""Device-side batch plan and operator calls (the host half of the C ABI).

`BatchPlan` turns (species, per-atom parameters) into the index tensors of `seqm_batch_t`; the `op_*`
functions are 1:1 wrappers of the C entry points and are what pyseqm_b200/seqm_functions/* expose under
the reference's operator names.  Torch is used only for allocation, index arithmetic and streams.
"""
import ctypes as C
import json
import os

import torch

from ._lib import (JACOBI_NP, METHOD_ID, N_ELEM_ROWS, NPAR, PAR_ROWS, SeqmBatchStruct, SeqmError, SeqmPlanCounts,
                   SeqmScfOpts, ptr, stream_of)  # fmt: skip

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_TABLE_CACHE = {}


def element_tables():
    if "el" not in _TABLE_CACHE:
        with open(os.path.join(_DATA, "element_tables.json")) as f:
            _TABLE_CACHE["el"] = json.load(f)
    return _TABLE_CACHE["el"]


def method_table(method):
    """(zmax+1, ncols) float64 CPU tensor + column names (seqm/seqm_functions/parameters.py:4-46)."""
    key = ("par", method)
    if key not in _TABLE_CACHE:
        with open(os.path.join(_DATA, f"params_{method}.json")) as f:
            d = json.load(f)
        zmax = max(int(k) for k in d["rows"])
        tab = torch.zeros((zmax + 1, len(d["columns"])), dtype=torch.float64)
        for k, v in d["rows"].items():
            tab[int(k)] = torch.tensor(v, dtype=torch.float64)
        pw = None
        if "pairwise_alpha_chi" in d:  # PWCCT (parameters.py:49-88): alpha[Zi, Zj], chi[Zi, Zj]
            zm = max(max(t[0], t[1]) for t in d["pairwise_alpha_chi"])
            a = torch.zeros((zm + 1, zm + 1), dtype=torch.float64)
            c = torch.zeros_like(a)
            for zi, zj, al, ch in d["pairwise_alpha_chi"]:
                a[zi, zj] = al
                c[zi, zj] = ch
            pw = (a, c)
        _TABLE_CACHE[key] = (tab, d["columns"], pw)
    return _TABLE_CACHE[key]


def check_input(species):
    """Rows must be non-increasing in Z (Molecule.py:188-206, same message)."""
    ok = species[:, :-1] >= species[:, 1:]
    row_ok = ok.all(dim=1)
    if not bool(row_ok.all()):
        bad = (~row_ok).nonzero(as_tuple=False).squeeze(1).tolist()
        rows = ", ".join(map(str, bad))
        row_word = "row" if len(bad) == 1 else "rows"
        verb = "is" if len(bad) == 1 else "are"
        raise ValueError(f"species must be non-increasing along each row, but {row_word} {rows} {verb} not sorted.")


def _pm6d_host_tables():
    """Element-level tables of the PM6 d-orbital path (seqm_functions/pm6d_tables.py), built once per process:
    d rows of the element table, one-centre d integrals per element, multipole coefficients, overlap polynomials."""
    if "pm6d" not in _TABLE_CACHE:
        from .seqm_functions import pm6d_tables as T

        tab, cols, _ = method_table("PM6")
        el = element_tables()
        nz = max(tab.shape[0], len(el["tore"]))
        drows = torch.zeros((len(T.D_ROWS), nz), dtype=torch.float64)
        onec = torch.zeros((nz, 45 * 45), dtype=torch.float64)
        supported = set()
        for z in range(1, min(tab.shape[0], len(el["qn_int"]))):
            row = {c: float(tab[z, j]) for j, c in enumerate(cols)}
            if not any(row.values()) or row["zeta_s"] <= 0.0:
                continue
            qn, qnd = int(el["qn_int"][z]), int(el["qnD_int"][z])
            d = T.d_rows(z, qn, qnd, row)
            for r, name in enumerate(T.D_ROWS):
                drows[r, z] = d[name]
            if T.d_shell(z) and row["zeta_d"] != 0.0 and row["rho_core"] == 0.0 and qnd > 0:
                onec[z] = torch.as_tensor(T.one_center_integrals(qn, qnd, row["s_orb_exp_tail"], row["p_orb_exp_tail"],
                                                                 row["d_orb_exp_tail"], row["F0SD"], row["G2SD"]).reshape(-1))
                supported.add(z)
        c, cyx = T.multipole_coefficients()
        _TABLE_CACHE["pm6d"] = dict(drows=drows, onecenter=onec, supported=supported, mp_coef=torch.as_tensor(c).reshape(-1),
                                    mp_coef_yx=torch.as_tensor(cyx).reshape(-1),
                                    ovl_poly=torch.as_tensor(T.overlap_polynomials()).reshape(-1))  # fmt: skip
    return _TABLE_CACHE["pm6d"]


def _device_tables(method, dev):
    """Per-element tables resident on `dev` (cached): the rows of `atom_par` that come from the element (NPAR, zmax+1),
    tore, class bounds; method "PM6_D" adds the d-shell rows and the constant tables of the spd kernels."""
    key = ("dev", method, str(dev))
    if key not in _TABLE_CACHE:
        tab, cols, pw = method_table("PM6" if method == "PM6_D" else method)
        el = element_tables()
        nz = max(tab.shape[0], len(el["tore"]))
        rows = torch.zeros((NPAR, nz), dtype=torch.float64)
        for r, name in enumerate(PAR_ROWS[:24]):
            if name in cols:
                rows[r, : tab.shape[0]] = tab[:, cols.index(name)]
        rows[24, : len(el["tore"])] = torch.tensor(el["tore"], dtype=torch.float64)
        rows[25, : len(el["qn"])] = torch.tensor(el["qn"], dtype=torch.float64)
        if "rho_core" in cols:
            rows[26, : tab.shape[0]] = tab[:, cols.index("rho_core")]
        rows[27, : len(el["atomic_num"])] = torch.tensor(el["atomic_num"], dtype=torch.float64)
        extra = {}
        if method == "PM6_D":
            h = _pm6d_host_tables()
            rows[PAR_ROWS.index("U_dd") :, : h["drows"].shape[1]] = h["drows"]
            extra = {k: h[k].to(dev).contiguous() for k in ("onecenter", "mp_coef", "mp_coef_yx", "ovl_poly")}
            extra["supported_d"] = h["supported"]
        T = {
            "rows": rows.to(dev).contiguous(),
            "tore": rows[24].to(dev).contiguous(),
            "jacobi_bounds": torch.tensor([2 * q for q in JACOBI_NP], device=dev),
            "cls_ids": torch.arange(len(JACOBI_NP) + 1, device=dev).unsqueeze(0),
            "ones": torch.ones(1, dtype=torch.int64, device=dev),
        }
        T.update(extra)
        _TABLE_CACHE[key] = (T, cols, pw)
    return _TABLE_CACHE[key]


class BatchPlan:
    """Index tensors of one molecule batch (topology only; coordinates are passed per call)."""

    def __init__(self, lib, species, method, parameters=None, charges=0, table=None, outer_cutoff=None):
        if method not in METHOD_ID:
            raise NotImplementedError(
                f"method {method!r} is not implemented by the B200 path (supported: {sorted(METHOD_ID)})"
            )
        self.lib = lib
        dev = species.device
        self.device = dev
        nmol, molsize = species.shape
        self.nmol, self.molsize, self.method = nmol, molsize, method
        self.d_mode = method == "PM6_D"  # method="PM6" with d-shell elements: 9 orbitals on those atoms
        self.stride = 9 if self.d_mode else 4  # orbital slots per atom in the dense API layout
        T, cols, pw = _device_tables("PM6_D" if self.d_mode else (table or method), dev)
        sp64 = species.to(torch.int64).contiguous()
        ch = None
        if torch.is_tensor(charges):
            ch = charges.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        elif int(charges) != 0:
            ch = torch.full((nmol,), int(charges), dtype=torch.int64, device=dev)
        # ---- step 1 (seqm_plan_count): per-molecule arrays, scans, processing order, host scalars ------------
        mi = torch.empty(9 * nmol + 2 + 256, dtype=torch.int32, device=dev)  # one allocation, sliced below
        atom0, pair0 = mi[: nmol + 1], mi[nmol + 1 : 2 * nmol + 2]
        o = 2 * nmol + 2
        nheavy32, nhyd32, nocc32, order = (mi[o + k * nmol : o + (k + 1) * nmol] for k in range(4))
        cls_pair0 = mi[o + 4 * nmol : o + 7 * nmol]
        counts_dev = mi[o + 7 * nmol + (o + 7 * nmol) % 2 :]  # 8-byte aligned tail (>= sizeof(seqm_plan_counts_t))
        mat0 = torch.empty(nmol + 1, dtype=torch.int64, device=dev)
        nsh32 = torch.zeros(nmol, dtype=torch.int32, device=dev) if self.d_mode else None
        cnt = SeqmPlanCounts()
        st = stream_of(mi)
        nz = T["rows"].shape[1]
        lib.check(lib.dll.seqm_plan_count(ptr(sp64), nmol, molsize, ptr(ch), ptr(T["rows"]), nz, ptr(atom0), ptr(pair0),
                                          ptr(mat0), ptr(nheavy32), ptr(nhyd32), ptr(nocc32), ptr(order), ptr(cls_pair0),
                                          ptr(counts_dev), C.byref(cnt), ptr(nsh32), st), "seqm_plan_count")  # fmt: skip
        self.nat, self.npairs, self.mat_total, self.nmax = cnt.nat, cnt.npairs, cnt.mat_total, cnt.nmax
        self.zmax, self.sorted_ok = cnt.zmax, not cnt.unsorted
        self.elements = [0] + [z for z in range(1, 128) if cnt.elements[z]]
        if not self.sorted_ok:
            check_input(species)
            # rows are sorted by Z, so the flag came from the d-shell ordering test of the PM6 plan
            raise ValueError("method 'PM6': the d-shell elements of a molecule must precede its sp-only heavy elements in the "
                             "Z-sorted order (the reference's packd layout, packd.py:195-218, cannot hold e.g. Ca before Cl)")
        if cnt.odd_electrons:
            raise ValueError("RHF setting requires closed shell systems (even number of electrons)")
        # ---- step 2 (seqm_plan_fill): atoms, per-atom parameters, pair list, class-sorted pair ids -----------
        ai = torch.empty(2 * self.nat + 3 * self.npairs, dtype=torch.int32, device=dev)
        atom_Z32, atom_mol32 = ai[: self.nat], ai[self.nat : 2 * self.nat]
        pair_i32, pair_j32, self.pair_perm = (ai[2 * self.nat + k * self.npairs : 2 * self.nat + (k + 1) * self.npairs] for k in range(3))
        self.real_atoms = torch.empty(self.nat, dtype=torch.int64, device=dev)
        par = torch.zeros((NPAR, self.nat), dtype=torch.float64, device=dev)
        lib.check(lib.dll.seqm_plan_fill(ptr(sp64), nmol, molsize, C.byref(cnt), ptr(atom0), ptr(pair0), ptr(nheavy32),
                                         ptr(cls_pair0), ptr(T["rows"]), NPAR, nz, ptr(atom_Z32), ptr(atom_mol32),
                                         ptr(self.real_atoms), ptr(par), ptr(pair_i32), ptr(pair_j32), ptr(self.pair_perm),
                                         st), "seqm_plan_fill")  # fmt: skip
        self._keep = (mi, ai, sp64, ch, nsh32)
        self.real_mask = None
        self.t = dict(mol_atom0=atom0, mol_pair0=pair0, mol_mat0=mat0, mol_nheavy=nheavy32, mol_nhyd=nhyd32,
                      mol_nocc=nocc32, mol_order=order, atom_Z=atom_Z32, atom_mol=atom_mol32, pair_i=pair_i32,
                      pair_j=pair_j32)  # fmt: skip
        # int64 views for torch-side index arithmetic are made on first use (see __getattr__)
        pair_cls_cnt = list(cnt.pair_cls_cnt)
        cls_cnt = list(cnt.jacobi_cls_cnt)
        fock_scratch = cnt.fock_scratch
        ncls = len(JACOBI_NP)
        zmax = cnt.zmax
        # ---- per-atom parameter overrides (values only) ---------------------------------------------------
        if parameters:
            for r, name in enumerate(PAR_ROWS[:24]):
                if parameters.get(name) is not None:
                    par[r] = parameters[name].to(torch.float64)
        self.pw = None
        if pw is not None:
            dim = max(zmax + 1, 2)
            key = ("pwdev", table or method, str(dev), dim)
            if key not in _TABLE_CACHE:
                a = torch.zeros((dim, dim), dtype=torch.float64)
                c = torch.zeros_like(a)
                k = min(dim, pw[0].shape[0])
                a[:k, :k] = pw[0][:k, :k]
                c[:k, :k] = pw[1][:k, :k]
                _TABLE_CACHE[key] = (a.to(dev).contiguous(), c.to(dev).contiguous(), dim)
            self.pw = _TABLE_CACHE[key]
        self.par = par
        self.par_version = 0  # bumped by set_parameters(): invalidates per-plan caches of parameter-derived sums
        s = SeqmBatchStruct()
        s.nmol, s.nat, s.npairs, s.method = nmol, self.nat, self.npairs, METHOD_ID[method]
        s.nmax, s.molsize, s.mat_total = self.nmax, molsize, self.mat_total
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        s.atom_par = self.par.data_ptr()
        if self.pw is not None:
            s.pw_alpha, s.pw_chi, s.pw_dim = self.pw[0].data_ptr(), self.pw[1].data_ptr(), self.pw[2]
        s.pair_cls_off[0], s.pair_cls_off[1] = 0, pair_cls_cnt[0]
        s.pair_cls_off[2], s.pair_cls_off[3] = pair_cls_cnt[0] + pair_cls_cnt[1], sum(pair_cls_cnt)
        s.pair_perm = self.pair_perm.data_ptr()
        s.fock_scratch = fock_scratch
        # Parser's pair_outer_cutoff in Angstrom (basics.py:209, 326); the reference default 1e10 keeps every pair
        s.pair_outer_cutoff = float(outer_cutoff) if outer_cutoff is not None and outer_cutoff < 1.0e9 else 0.0
        # eigensolver size classes over the descending-n processing order (host arrays inside the struct):
        # class c = smallest NP with 2*NP >= n; molecules beyond the last class (large path) belong to none
        begin = cls_cnt[ncls]  # mol_order is descending in n: the too-large molecules come first
        for c in range(ncls - 1, -1, -1):
            s.cls_begin[c] = begin
            s.cls_count[c] = cls_cnt[c]
            begin += cls_cnt[c]
        if self.d_mode:
            self._init_d_mode(s, T, nsh32)
        self.struct = s
        self.ref = C.byref(s)
        self.large = self.nmax > lib.dll.seqm_max_orbitals()  # global-memory Fock + GEMM SP2/DIIS path
        # the eigensolver route of the large path: one-sided Jacobi up to seqm_max_orbitals_eig() (256) orbitals
        self.eig_ok = self.nmax <= lib.dll.seqm_max_orbitals_eig()
        if self.d_mode and self.large:
            raise NotImplementedError(f"method 'PM6' with d orbitals: a molecule with {self.nmax} orbitals exceeds the "
                                      f"shared-memory resident path ({lib.dll.seqm_max_orbitals()} orbitals)")
        lib.check(lib.dll.seqm_atom_multipoles(self.ref, stream_of(par)), "seqm_atom_multipoles")

    def _init_d_mode(self, s, T, nsh32):
        """Index arrays of the pairs that contain a d atom (torch index arithmetic, once per plan): ragged offsets of
        their (np_i x np_j) integral blocks, the list of those pairs, and the constant tables of the spd kernels."""
        dev = self.device
        bad = sorted({int(z) for z in self.elements if z and _is_d_shell(z)} - T["supported_d"])
        if bad:
            raise NotImplementedError(f"method 'PM6': no usable d-orbital parameters for Z={bad} (no zeta_d in the parameter "
                                      "file, or a rho_core element)")
        elq = element_tables()["qn_int"]
        big = sorted(int(z) for z in self.elements if z and not _is_d_shell(z) and z > 1 and elq[z] > 3)
        if big:
            raise ValueError("\nError from diat.py, overlap matrix\nSome elements are not supported yet")
        nsh = nsh32.to(torch.int64)
        mol = self.t["atom_mol"].to(torch.int64)
        local = torch.arange(self.nat, device=dev) - self.t["mol_atom0"].to(torch.int64)[mol]
        Z = self.t["atom_Z"].to(torch.int64)
        is_d = local < nsh[mol]
        nprod = torch.where(is_d, 45, torch.where(Z > 1, 10, 1))
        pi, pj = self.t["pair_i"].to(torch.int64), self.t["pair_j"].to(torch.int64)
        size = torch.where(is_d[pi], nprod[pi] * nprod[pj], 0)
        wd0 = torch.zeros(self.npairs + 1, dtype=torch.int64, device=dev)
        torch.cumsum(size, 0, out=wd0[1:])
        yp = torch.nonzero(size > 0, as_tuple=False).squeeze(1)
        slot = torch.full((max(self.npairs, 1),), -1, dtype=torch.int32, device=dev)
        slot[yp] = torch.arange(yp.numel(), dtype=torch.int32, device=dev)
        self.nsh = nsh
        self.n_ypairs = int(yp.numel())
        self.wd_total = int(wd0[-1])
        self.pair_wd0, self.ypairs, self.ypair_slot = wd0, yp.to(torch.int32).contiguous(), slot
        self.pair_nprod = (nprod[pi], nprod[pj])
        self._dkeep = (T["onecenter"], T["mp_coef"], T["mp_coef_yx"], T["ovl_poly"])
        s.mol_nsh = nsh32.data_ptr()
        s.pair_wd0, s.ypairs, s.ypair_slot = wd0.data_ptr(), self.ypairs.data_ptr(), slot.data_ptr()
        s.n_ypairs, s.oc_dim = self.n_ypairs, T["onecenter"].shape[0]
        s.onecenter_d, s.mp_coef = T["onecenter"].data_ptr(), T["mp_coef"].data_ptr()
        s.mp_coef_yx, s.ovl_poly = T["mp_coef_yx"].data_ptr(), T["ovl_poly"].data_ptr()

    _LAZY64 = {"nheavy": "mol_nheavy", "nhyd": "mol_nhyd", "nocc": "mol_nocc", "Z": "atom_Z", "atom_mol": "atom_mol",
               "pair_i": "pair_i", "pair_j": "pair_j"}  # fmt: skip

    def __getattr__(self, name):  # int64 copies of the int32 kernel arrays, for torch indexing, on first use
        d = self.__dict__
        if name in BatchPlan._LAZY64 and "t" in d:
            d[name] = d["t"][BatchPlan._LAZY64[name]].to(torch.int64)
        elif name == "na" and "t" in d:
            d["na"] = self.nheavy + self.nhyd
        elif name == "norb" and "t" in d:
            d["norb"] = 4 * self.nheavy + self.nhyd + (5 * d["nsh"] if d.get("d_mode") else 0)
        elif name == "atom_local" and "t" in d:
            d["atom_local"] = torch.arange(self.nat, device=self.device) - d["t"]["mol_atom0"].to(torch.int64)[self.atom_mol]
        else:
            raise AttributeError(name)
        return d[name]

    def set_parameters(self, learned):
        """Overwrite per-atom parameter rows (values only) and rebuild the multipole rows that depend on them."""
        if not learned:
            return
        for name, t in learned.items():
            if t.requires_grad:
                raise NotImplementedError("gradients with respect to learned parameters are not on the B200 path")
            self.par[PAR_ROWS.index(name)] = t.detach().to(torch.float64)
        self.par_version += 1
        self.lib.check(self.lib.dll.seqm_atom_multipoles(self.ref, stream_of(self.par)), "seqm_atom_multipoles")

    # ---- helpers -----------------------------------------------------------------------------------
    def new_mat(self):
        return torch.zeros(self.mat_total, dtype=torch.float64, device=self.device)

    def real_xyz(self, coordinates):
        return coordinates.detach().reshape(-1, 3)[self.real_atoms].contiguous()

    def parameter(self, name):
        return self.par[PAR_ROWS.index(name)]


def _is_d_shell(z):
    return (12 < z < 18) or (20 < z < 30) or (32 < z < 36) or (38 < z < 48) or (50 < z < 54) or (70 < z < 80) or z == 57


def op_pair_integrals(plan, xyz):
    w = torch.empty((plan.npairs, 10, 10), dtype=torch.float64, device=plan.device)
    hab = torch.empty((plan.npairs, 4, 4), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_pair_integrals(plan.ref, ptr(xyz), ptr(w), ptr(hab), stream_of(xyz)), "seqm_pair_integrals")
    if plan.d_mode:
        op_pair_integrals_d(plan, xyz, w)
    return w, hab


def op_pair_integrals_d(plan, xyz, w):
    """PM6 with d orbitals: the ragged integral blocks and 9 x 9 overlap blocks of the pairs with a d atom.  They belong
    to the current geometry and are attached to the plan's batch struct (b->wd, b->hab_d) for the kernels that follow."""
    wd = torch.empty(max(plan.wd_total, 1), dtype=torch.float64, device=plan.device)
    hab_d = torch.empty((max(plan.n_ypairs, 1), 9, 9), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_pair_integrals_d(plan.ref, ptr(xyz), ptr(w), ptr(wd), ptr(hab_d), stream_of(xyz)),
                   "seqm_pair_integrals_d")
    plan._wd = (wd, hab_d)  # keeps the buffers alive as long as the struct points at them
    plan.struct.wd, plan.struct.hab_d = wd.data_ptr(), hab_d.data_ptr()
    return wd, hab_d


def dense_w45(plan, w, wd):
    """The reference's `molecule.w` for method="PM6": (npairs, 45, 45) with the roles of the two atoms swapped
    (w[p, mn on j, kl on i]; hcore.py:143-146, fock.py:277), assembled from the dense sp blocks and the ragged d blocks."""
    out = torch.zeros((plan.npairs, 45, 45), dtype=torch.float64, device=plan.device)
    out[:, :10, :10] = w.transpose(1, 2)
    if plan.d_mode and plan.n_ypairs:
        yp = plan.ypairs.to(torch.int64)
        npj = plan.pair_nprod[1][yp]
        off = plan.pair_wd0[yp]
        for n in (45, 10, 1):
            sel = torch.nonzero(npj == n, as_tuple=False).squeeze(1)
            if sel.numel() == 0:
                continue
            idx = off[sel].unsqueeze(1) + torch.arange(45 * n, device=plan.device).unsqueeze(0)
            out[yp[sel], :n, :] = wd[idx].reshape(-1, 45, n).transpose(1, 2)
    return out


def op_hcore(plan, w, hab):
    H = plan.new_mat()
    plan.lib.check(plan.lib.dll.seqm_hcore(plan.ref, ptr(w), ptr(hab), ptr(H), stream_of(H)), "seqm_hcore")
    return H


def op_fock(plan, P, H, w, active=None, out=None):
    F = plan.new_mat() if out is None else out
    plan.lib.check(plan.lib.dll.seqm_fock(plan.ref, ptr(P), ptr(H), ptr(w), ptr(F), ptr(active), stream_of(F)), "seqm_fock")
    return F


def op_eig_density(plan, F, want_P=True, want_C=False, Cguess=None, active=None, want_e=True):
    """want_e=False asks for the density only: the eigensolver may then finish with its first-order
    occupied-virtual correction instead of a last sweep (eig_kernels.cuh); eigenvalues are not returned."""
    if Cguess is not None and (Cguess.numel() != plan.mat_total or Cguess.device != plan.device):
        raise SeqmError(f"op_eig_density: warm-start eigenvectors hold {Cguess.numel()} elements on {Cguess.device}, the "
                        f"plan has {plan.mat_total} on {plan.device} (a guess from another batch plan?)")
    if plan.large:  # mid-size eigensolver (one-sided Jacobi): works in the eigenvector buffer, needs both; always cold
        want_P = want_C = True
        Cguess = None
    P = plan.new_mat() if (want_P or Cguess is not None) else None
    Cm = plan.new_mat() if (want_C or Cguess is not None) else None  # the warm start needs both scratch slots
    e = torch.zeros((plan.nmol, plan.nmax), dtype=torch.float64, device=plan.device) if want_e else None
    plan.lib.check(
        plan.lib.dll.seqm_eig_density(plan.ref, ptr(F), ptr(P), ptr(e), ptr(Cm), ptr(Cguess), ptr(active), stream_of(F)),
        "seqm_eig_density",
    )
    return e, P, Cm


def op_sp2_density(plan, F, eps, active=None):
    P = plan.new_mat()
    if plan.large:
        nb = plan.lib.dll.seqm_sp2_large_workspace_bytes(plan.ref)
        ws = torch.zeros(nb, dtype=torch.uint8, device=plan.device)
        nit_h = (C.c_int32 * plan.nmol)()
        plan.lib.check(
            plan.lib.dll.seqm_sp2_density_large(plan.ref, ptr(F), ptr(P), C.c_double(eps), nit_h, ptr(ws), stream_of(F)),
            "seqm_sp2_density_large",
        )
        return P, torch.tensor(list(nit_h), dtype=torch.int32)
    nit = torch.zeros(plan.nmol, dtype=torch.int32, device=plan.device)
    plan.lib.check(
        plan.lib.dll.seqm_sp2_density(plan.ref, ptr(F), ptr(P), C.c_double(eps), ptr(nit), ptr(active), stream_of(F)),
        "seqm_sp2_density",
    )
    return P, nit


def op_elec_energy(plan, P, H, F):
    E = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_elec_energy(plan.ref, ptr(P), ptr(H), ptr(F), ptr(E), None, stream_of(E)), "seqm_elec_energy")
    return E


def op_nuclear_energy(plan, xyz, w):
    EAB = torch.zeros(max(plan.npairs, 1), dtype=torch.float64, device=plan.device)
    En = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_nuclear_energy(plan.ref, ptr(xyz), ptr(w), ptr(EAB), ptr(En), stream_of(En)), "seqm_nuclear_energy")
    return EAB[: plan.npairs], En


def op_gradient(plan, xyz, P, forward_mode=False):
    scratch = torch.zeros((max(plan.npairs, 1), 3), dtype=torch.float64, device=plan.device)
    g = torch.zeros((plan.nat, 3), dtype=torch.float64, device=plan.device)
    fn = plan.lib.dll.seqm_gradient_forward if forward_mode else plan.lib.dll.seqm_gradient
    plan.lib.check(fn(plan.ref, ptr(xyz), ptr(P), ptr(scratch), ptr(g), stream_of(g)), "seqm_gradient")
    return g


def op_gradient_xl(plan, xyz, D, P):
    scratch = torch.zeros((max(plan.npairs, 1), 3), dtype=torch.float64, device=plan.device)
    g = torch.zeros((plan.nat, 3), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_gradient_xl(plan.ref, ptr(xyz), ptr(D), ptr(P), ptr(scratch), ptr(g), stream_of(g)), "seqm_gradient_xl")
    return g


def op_elec_energy_xl(plan, D, P, F, H):
    E = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_elec_energy_xl(plan.ref, ptr(D), ptr(P), ptr(F), ptr(H), ptr(E), stream_of(E)), "seqm_elec_energy_xl")
    return E


def op_xl_propagate(plan, kappa, c, D, P, Pt, coef, slot):
    """P(n+1) = kappa [c D + (1-c) P] + sum_j coef_j Pt_j, stored into Pt[slot] as well; returns the new field density."""
    out = torch.empty_like(P)
    plan.lib.check(plan.lib.dll.seqm_xl_propagate(P.numel(), float(kappa), float(c), ptr(D), ptr(P), ptr(Pt), ptr(coef),
                                                  Pt.shape[0], int(slot), ptr(out), stream_of(P)), "seqm_xl_propagate")  # fmt: skip
    return out


def op_orbitals_dense(plan, Cm):
    V = torch.empty((plan.nmol, plan.nmax, plan.nmax), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_orbitals_dense(plan.ref, ptr(Cm), ptr(V), stream_of(V)), "seqm_orbitals_dense")
    return V


def op_post_scf(plan, P, xyz, g=None, want_dipole=True):
    """Mulliken charges (nmol, molsize), ground-state dipole (nmol, 3) and force = -g in the padded (nmol, molsize, 3)
    layout, one launch on the packed density (seqm_post_scf).  g: (nat, 3) from op_gradient, or None."""
    from .seqm_functions.constants import a0, debye_to_AU, to_debye

    q = torch.empty((plan.nmol, plan.molsize), dtype=torch.float64, device=plan.device)
    dip = torch.empty((plan.nmol, 3), dtype=torch.float64, device=plan.device) if want_dipole else None
    force = torch.empty((plan.nmol, plan.molsize, 3), dtype=torch.float64, device=plan.device) if g is not None else None
    if g is not None:
        g = g.contiguous()
    plan.lib.check(plan.lib.dll.seqm_post_scf(plan.ref, ptr(P), ptr(xyz), ptr(g), ptr(q), ptr(dip), ptr(force), float(a0),
                                              float(to_debye * debye_to_AU), stream_of(q)), "seqm_post_scf")  # fmt: skip
    return q, dip, force


# ---- KSA-XL-BOMD building blocks (seqm_ksa.cu) --------------------------------------------------------------------------
def op_packed_gemm(plan, A, B, ta=False, tb=False):
    out = plan.new_mat()
    plan.lib.check(plan.lib.dll.seqm_packed_gemm(plan.ref, ptr(A), ptr(B), ptr(out), int(ta), int(tb), stream_of(out)),
                   "seqm_packed_gemm")  # fmt: skip
    return out


def op_scale_columns(plan, Cm, f, s):
    out = plan.new_mat()
    f = f.contiguous()
    plan.lib.check(plan.lib.dll.seqm_scale_columns(plan.ref, ptr(Cm), ptr(f), float(s), ptr(out), stream_of(out)), "seqm_scale_columns")
    return out


def op_canon_prt(plan, e, mu, X, beta, m_iter):
    """in place on X (packed, eigenbasis)"""
    e, mu = e.contiguous(), mu.contiguous()
    plan.lib.check(plan.lib.dll.seqm_canon_prt(plan.ref, ptr(e), ptr(mu), ptr(X), float(beta), int(m_iter), stream_of(X)), "seqm_canon_prt")
    return X


def op_packed_dot(plan, X, Y):
    out = torch.empty((plan.nmol,), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_packed_dot(plan.ref, ptr(X), ptr(Y), ptr(out), stream_of(out)), "seqm_packed_dot")
    return out


def op_packed_axpby(plan, a, X, c, Y):
    """Y_m = a[m] X_m + c[m] Y_m in place (a / c None: 1; X None: pure scaling)"""
    a = None if a is None else a.contiguous()
    c = None if c is None else c.contiguous()
    plan.lib.check(plan.lib.dll.seqm_packed_axpby(plan.ref, ptr(a), ptr(X), ptr(c), ptr(Y), stream_of(Y)), "seqm_packed_axpby")
    return Y


def op_sigma_ao(plan, X, w, all_symmetric=False):
    """Two-electron response F[X] of a (generally non-symmetric) packed AO matrix X, e.g. a CIS transition density
    (makeA_pi_batched, rcis_batch.py:296-403): the Fock build of its symmetric part without Hcore plus the exchange-only
    response of its antisymmetric part."""
    XT = torch.empty_like(X)
    plan.lib.check(plan.lib.dll.seqm_packed_transpose(plan.ref, ptr(X), ptr(XT), stream_of(X)), "seqm_packed_transpose")
    F = op_fock(plan, 0.5 * (X + XT), plan.new_mat(), w)
    if not all_symmetric:
        Fa = torch.empty_like(X)
        Xa = 0.5 * (X - XT)
        plan.lib.check(plan.lib.dll.seqm_fock_antisym(plan.ref, ptr(Xa), ptr(w), ptr(Fa), stream_of(Fa)), "seqm_fock_antisym")
        F += Fa
    return F


def op_mo_match(plan, V_new, V_old, e):
    """Energy._crossing_match_molecular_orbitals[_grouped] (basics.py:596-719): V (nmol, nmax, nmax), e (nmol, nmax)."""
    V_new, V_old, e = V_new.contiguous(), V_old.contiguous(), e.contiguous()
    S = plan.new_mat()
    ints = torch.empty((2, plan.nmol, plan.nmax), dtype=torch.int32, device=plan.device)
    prio = torch.empty((plan.nmol, plan.nmax), dtype=torch.float64, device=plan.device)
    V_out, e_out = torch.empty_like(V_new), torch.empty_like(e)
    plan.lib.check(plan.lib.dll.seqm_mo_match(plan.ref, ptr(V_new), ptr(V_old), ptr(e), ptr(S), ptr(ints[0]), ptr(ints[1]),
                                              ptr(prio), ptr(V_out), ptr(e_out), stream_of(V_new)), "seqm_mo_match")  # fmt: skip
    return V_out, e_out


def op_pack(plan, dense):
    out = plan.new_mat()
    d = dense.detach().contiguous()
    plan.lib.check(plan.lib.dll.seqm_pack(plan.ref, ptr(d), ptr(out), stream_of(out)), "seqm_pack")
    return out


def op_unpack(plan, packed, out=None):
    N = plan.stride * plan.molsize
    if out is None:
        out = torch.empty((plan.nmol, N, N), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_unpack(plan.ref, ptr(packed), ptr(out), stream_of(out)), "seqm_unpack")
    return out


def op_initial_density(plan):
    P = plan.new_mat()
    plan.lib.check(plan.lib.dll.seqm_initial_density(plan.ref, ptr(P), stream_of(P)), "seqm_initial_density")
    return P


def op_scf(plan, H, w, P, eps, converger, sp2=(False,), max_iter=1000, warm_start=True, want_C=False, C0=None):
    """Runs the SCF loop; P (packed) is updated in place.  Returns (F, Eelec, notconverged, n_iter[, C]).
    C0: packed eigenvectors of a nearby problem on the same plan; the first density solve starts from them."""
    o = SeqmScfOpts()
    o.eps = float(eps)
    o.converger = int(converger[0])
    o.alpha = float(converger[1]) if (o.converger == 0 and len(converger) > 1) else 0.0
    o.use_sp2 = 1 if sp2[0] else 0
    o.sp2_eps = float(sp2[1]) if sp2[0] else 0.0
    o.max_iter = int(max_iter)
    o.warm_start = (2 if (C0 is not None and want_C and not sp2[0]) else 1) if warm_start else 0
    o.pipeline = int(os.environ.get("SEQM_B200_PIPELINE", "0"))  # 0 auto, 1 single stream, 2 two half-batches
    nbytes = plan.lib.dll.seqm_scf_workspace_bytes(plan.ref, C.byref(o))
    if nbytes < 0:
        raise SeqmError("seqm_scf_workspace_bytes failed")
    # the workspace (DIIS history ~0.8 GB at configs[1]) belongs to the plan: allocated and zeroed once, reused by every
    # later forward -- scf_init_kernel resets all the state the loop reads before writing
    ws = plan.__dict__.get("_scf_ws")
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=plan.device)
        plan.__dict__["_scf_ws"] = ws
    F = plan.new_mat()
    E = torch.zeros(plan.nmol, dtype=torch.float64, device=plan.device)
    nc = torch.ones(plan.nmol, dtype=torch.int32, device=plan.device)
    nit = C.c_int32(0)
    Clast = plan.new_mat() if (want_C and not sp2[0]) else None
    if o.warm_start == 2:
        Clast.copy_(C0)
    plan.lib.check(
        plan.lib.dll.seqm_scf(plan.ref, C.byref(o), ptr(H), ptr(w), ptr(P), ptr(F), ptr(E), ptr(nc), ptr(ws),
                              C.byref(nit), ptr(Clast), stream_of(P)),
        "seqm_scf",
    )  # fmt: skip
    if want_C:
        return F, E, nc.to(torch.bool), int(nit.value), Clast
    return F, E, nc.to(torch.bool), int(nit.value)
