"""Device-side batch plan and operator calls (the host half of the C ABI).

`BatchPlan` turns (species, per-atom parameters) into the index tensors of `seqm_batch_t`; the `op_*`
functions are 1:1 wrappers of the C entry points and are what pyseqm_b200/seqm_functions/* expose under
the reference's operator names.  Torch is used only for allocation, index arithmetic and streams.
"""
import ctypes as C
import json
import os

import torch

from ._lib import (JACOBI_NP, METHOD_ID, NPAR, PAR_ROWS, SeqmBatchStruct, SeqmError, SeqmPlanCounts, SeqmScfOpts, ptr,
                   stream_of)  # fmt: skip

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


def _device_tables(method, dev):
    """Per-element tables resident on `dev` (cached): rows 0..27 of `atom_par` as (28, zmax+1), tore, class bounds."""
    key = ("dev", method, str(dev))
    if key not in _TABLE_CACHE:
        tab, cols, pw = method_table(method)
        el = element_tables()
        nz = max(tab.shape[0], len(el["tore"]))
        rows = torch.zeros((28, nz), dtype=torch.float64)
        for r, name in enumerate(PAR_ROWS[:24]):
            if name in cols:
                rows[r, : tab.shape[0]] = tab[:, cols.index(name)]
        rows[24, : len(el["tore"])] = torch.tensor(el["tore"], dtype=torch.float64)
        rows[25, : len(el["qn"])] = torch.tensor(el["qn"], dtype=torch.float64)
        if "rho_core" in cols:
            rows[26, : tab.shape[0]] = tab[:, cols.index("rho_core")]
        rows[27, : len(el["atomic_num"])] = torch.tensor(el["atomic_num"], dtype=torch.float64)
        T = {
            "rows": rows.to(dev).contiguous(),
            "tore": rows[24].to(dev).contiguous(),
            "jacobi_bounds": torch.tensor([2 * q for q in JACOBI_NP], device=dev),
            "cls_ids": torch.arange(len(JACOBI_NP) + 1, device=dev).unsqueeze(0),
            "ones": torch.ones(1, dtype=torch.int64, device=dev),
        }
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
        T, cols, pw = _device_tables(table or method, dev)
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
        cnt = SeqmPlanCounts()
        st = stream_of(mi)
        nz = T["rows"].shape[1]
        lib.check(lib.dll.seqm_plan_count(ptr(sp64), nmol, molsize, ptr(ch), ptr(T["rows"]), nz, ptr(atom0), ptr(pair0),
                                          ptr(mat0), ptr(nheavy32), ptr(nhyd32), ptr(nocc32), ptr(order), ptr(cls_pair0),
                                          ptr(counts_dev), C.byref(cnt), st), "seqm_plan_count")  # fmt: skip
        self.nat, self.npairs, self.mat_total, self.nmax = cnt.nat, cnt.npairs, cnt.mat_total, cnt.nmax
        self.zmax, self.sorted_ok = cnt.zmax, not cnt.unsorted
        self.elements = [0] + [z for z in range(1, 128) if cnt.elements[z]]
        if not self.sorted_ok:
            check_input(species)
        if cnt.odd_electrons:
            raise ValueError("RHF setting requires closed shell systems (even number of electrons)")
        # ---- step 2 (seqm_plan_fill): atoms, per-atom parameters, pair list, class-sorted pair ids -----------
        ai = torch.empty(2 * self.nat + 3 * self.npairs, dtype=torch.int32, device=dev)
        atom_Z32, atom_mol32 = ai[: self.nat], ai[self.nat : 2 * self.nat]
        pair_i32, pair_j32, self.pair_perm = (ai[2 * self.nat + k * self.npairs : 2 * self.nat + (k + 1) * self.npairs] for k in range(3))
        self.real_atoms = torch.empty(self.nat, dtype=torch.int64, device=dev)
        par = torch.zeros((NPAR, self.nat), dtype=torch.float64, device=dev)
        lib.check(lib.dll.seqm_plan_fill(ptr(sp64), nmol, molsize, C.byref(cnt), ptr(atom0), ptr(pair0), ptr(nheavy32),
                                         ptr(cls_pair0), ptr(T["rows"]), 28, nz, ptr(atom_Z32), ptr(atom_mol32),
                                         ptr(self.real_atoms), ptr(par), ptr(pair_i32), ptr(pair_j32), ptr(self.pair_perm),
                                         st), "seqm_plan_fill")  # fmt: skip
        self._keep = (mi, ai, sp64, ch)
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
        self.struct = s
        self.ref = C.byref(s)
        self.large = self.nmax > lib.dll.seqm_max_orbitals()  # global-memory Fock + GEMM SP2/DIIS path
        lib.check(lib.dll.seqm_atom_multipoles(self.ref, stream_of(par)), "seqm_atom_multipoles")

    _LAZY64 = {"nheavy": "mol_nheavy", "nhyd": "mol_nhyd", "nocc": "mol_nocc", "Z": "atom_Z", "atom_mol": "atom_mol",
               "pair_i": "pair_i", "pair_j": "pair_j"}  # fmt: skip

    def __getattr__(self, name):  # int64 copies of the int32 kernel arrays, for torch indexing, on first use
        d = self.__dict__
        if name in BatchPlan._LAZY64 and "t" in d:
            d[name] = d["t"][BatchPlan._LAZY64[name]].to(torch.int64)
        elif name == "na" and "t" in d:
            d["na"] = self.nheavy + self.nhyd
        elif name == "norb" and "t" in d:
            d["norb"] = 4 * self.nheavy + self.nhyd
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


def op_pair_integrals(plan, xyz):
    w = torch.empty((plan.npairs, 10, 10), dtype=torch.float64, device=plan.device)
    hab = torch.empty((plan.npairs, 4, 4), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_pair_integrals(plan.ref, ptr(xyz), ptr(w), ptr(hab), stream_of(xyz)), "seqm_pair_integrals")
    return w, hab


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
    N = 4 * plan.molsize
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
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=plan.device)
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
