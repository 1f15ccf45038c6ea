"""Device-side batch plan and operator calls (the host half of the C ABI).

`BatchPlan` turns (species, per-atom parameters) into the index tensors of `seqm_batch_t`; the `op_*`
functions are 1:1 wrappers of the C entry points and are what pyseqm_b200/seqm_functions/* expose under
the reference's operator names.  Torch is used only for allocation, index arithmetic and streams.
"""
import ctypes as C
import json
import os

import torch

from ._lib import JACOBI_NP, METHOD_ID, NPAR, PAR_ROWS, SeqmBatchStruct, SeqmError, SeqmScfOpts, ptr, stream_of

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


class BatchPlan:
    """Index tensors of one molecule batch (topology only; coordinates are passed per call)."""

    def __init__(self, lib, species, method, parameters=None, charges=0, table=None):
        if method not in METHOD_ID:
            raise NotImplementedError(
                f"method {method!r} is not implemented by the B200 path (supported: {sorted(METHOD_ID)})"
            )
        self.lib = lib
        dev = species.device
        self.device = dev
        el = element_tables()
        nmol, molsize = species.shape
        self.nmol, self.molsize, self.method = nmol, molsize, method
        real = species > 0
        self.real_mask = real
        self.real_atoms = torch.nonzero(real.reshape(-1), as_tuple=False).squeeze(1)
        Z = species.reshape(-1)[self.real_atoms]
        self.Z = Z
        na = real.sum(dim=1)
        nheavy = (species > 1).sum(dim=1)
        nhyd = (species == 1).sum(dim=1)
        tore = torch.tensor(el["tore"], dtype=torch.float64, device=dev)
        nel = tore[species].sum(dim=1).to(torch.int64)
        if torch.is_tensor(charges):
            nel = nel - charges.reshape(-1).to(torch.int64).to(dev)
        else:
            nel = nel - int(charges)
        if bool(((nel % 2) == 1).any()):
            raise ValueError("RHF setting requires closed shell systems (even number of electrons)")
        nocc = nel // 2
        norb = 4 * nheavy + nhyd
        self.na, self.nheavy, self.nhyd, self.nocc, self.norb = na, nheavy, nhyd, nocc, norb
        self.nat = int(Z.shape[0])
        self.nmax = int(norb.max())
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        atom0 = torch.cat([zero, torch.cumsum(na, 0)])
        npair_m = na * (na - 1) // 2
        pair0 = torch.cat([zero, torch.cumsum(npair_m, 0)])
        nn = norb * norb
        nn = nn + (nn % 2)
        mat0 = torch.cat([zero, torch.cumsum(nn, 0)])
        self.mat_total = int(mat0[-1])
        self.npairs = int(pair0[-1])
        atom_mol = torch.repeat_interleave(torch.arange(nmol, device=dev), na)
        local = torch.arange(self.nat, device=dev) - atom0[atom_mol]
        # dense triangular pair list ordered (molecule, i, j)
        cnt = na[atom_mol] - 1 - local
        pair_i = torch.repeat_interleave(torch.arange(self.nat, device=dev), cnt)
        first_of_i = torch.cumsum(cnt, 0) - cnt
        pair_j = pair_i + 1 + (torch.arange(self.npairs, device=dev) - first_of_i[pair_i])
        order = torch.argsort(norb, descending=True, stable=True)
        i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
        self.t = dict(
            mol_atom0=i32(atom0), mol_pair0=i32(pair0), mol_mat0=mat0.contiguous(), mol_nheavy=i32(nheavy),
            mol_nhyd=i32(nhyd), mol_nocc=i32(nocc), mol_order=i32(order), atom_Z=i32(Z), atom_mol=i32(atom_mol),
            pair_i=i32(pair_i), pair_j=i32(pair_j),
        )  # fmt: skip
        self.atom_mol, self.atom_local = atom_mol, local
        self.pair_i, self.pair_j = pair_i, pair_j
        # per-atom parameter table
        tab, cols, pw = method_table(table or method)
        tabd = tab.to(dev)
        par = torch.zeros((NPAR, self.nat), dtype=torch.float64, device=dev)
        for r, name in enumerate(PAR_ROWS[:24]):
            if parameters is not None and name in parameters and parameters[name] is not None:
                par[r] = parameters[name].to(torch.float64)
            elif name in cols:
                par[r] = tabd[Z, cols.index(name)]
        par[24] = tore[Z]
        par[25] = torch.tensor(el["qn"], dtype=torch.float64, device=dev)[Z]
        if "rho_core" in cols:
            par[26] = tabd[Z, cols.index("rho_core")]
        par[27] = torch.tensor(el["atomic_num"], dtype=torch.float64, device=dev)[Z]
        self.pw = None
        if pw is not None:
            zmax = int(species.max())
            dim = max(zmax + 1, 2)
            a = torch.zeros((dim, dim), dtype=torch.float64)
            c = torch.zeros_like(a)
            k = min(dim, pw[0].shape[0])
            a[:k, :k] = pw[0][:k, :k]
            c[:k, :k] = pw[1][:k, :k]
            self.pw = (a.to(dev).contiguous(), c.to(dev).contiguous(), dim)
        self.par = par.contiguous()
        s = SeqmBatchStruct()
        s.nmol, s.nat, s.npairs, s.method = nmol, self.nat, self.npairs, METHOD_ID[method]
        s.nmax, s.molsize, s.mat_total = self.nmax, molsize, self.mat_total
        for k, v in self.t.items():
            setattr(s, k, v.data_ptr())
        s.atom_par = self.par.data_ptr()
        if self.pw is not None:
            s.pw_alpha, s.pw_chi, s.pw_dim = self.pw[0].data_ptr(), self.pw[1].data_ptr(), self.pw[2]
        # pair classes: 0 H-H, 1 X-H, 2 X-X (rows are sorted by descending Z, so Z_i >= Z_j for every pair)
        pcls = (Z[pair_i] > 1).to(torch.int64) + (Z[pair_j] > 1).to(torch.int64)
        self.pair_perm = torch.argsort(pcls, stable=True).to(torch.int32).contiguous()
        cnt = torch.bincount(pcls, minlength=3).cpu().tolist()
        s.pair_cls_off[0], s.pair_cls_off[1] = 0, cnt[0]
        s.pair_cls_off[2], s.pair_cls_off[3] = cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]
        s.pair_perm = self.pair_perm.data_ptr()
        nxx, nxh, nhh = nheavy * (nheavy - 1) // 2, nheavy * nhyd, nhyd * (nhyd - 1) // 2
        s.fock_scratch = int((20 * nxx + 11 * nxh + 2 * nhh).max())
        # eigensolver size classes over the descending-n processing order (host arrays inside the struct):
        # class c = smallest NP with 2*NP >= n; molecules beyond the last class (large path) belong to none
        bounds = torch.tensor([2 * q for q in JACOBI_NP], device=dev)
        cls = torch.bucketize(norb, bounds)  # == len(JACOBI_NP) for n > 2*NP_max
        cnt = torch.bincount(cls, minlength=len(JACOBI_NP) + 1).cpu().tolist()
        begin = cnt[len(JACOBI_NP)]  # mol_order is descending in n: the too-large molecules come first
        for c in range(len(JACOBI_NP) - 1, -1, -1):
            s.cls_begin[c] = begin
            s.cls_count[c] = cnt[c]
            begin += cnt[c]
        self.struct = s
        self.ref = C.byref(s)
        self.large = self.nmax > lib.dll.seqm_max_orbitals()  # global-memory Fock + GEMM SP2/DIIS path
        z = torch.zeros(1, dtype=torch.float64, device=dev)
        lib.check(lib.dll.seqm_atom_multipoles(self.ref, stream_of(z)), "seqm_atom_multipoles")

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


def op_eig_density(plan, F, want_P=True, want_C=False, Cguess=None, active=None):
    P = plan.new_mat() if (want_P or Cguess is not None) else None
    Cm = plan.new_mat() if (want_C or Cguess is not None) else None  # the warm start needs both scratch slots
    e = torch.zeros((plan.nmol, plan.nmax), dtype=torch.float64, device=plan.device)
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


def op_orbitals_dense(plan, Cm):
    V = torch.empty((plan.nmol, plan.nmax, plan.nmax), dtype=torch.float64, device=plan.device)
    plan.lib.check(plan.lib.dll.seqm_orbitals_dense(plan.ref, ptr(Cm), ptr(V), stream_of(V)), "seqm_orbitals_dense")
    return V


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


def op_scf(plan, H, w, P, eps, converger, sp2=(False,), max_iter=1000, warm_start=True, want_C=False):
    """Runs the SCF loop; P (packed) is updated in place.  Returns (F, Eelec, notconverged, n_iter)."""
    o = SeqmScfOpts()
    o.eps = float(eps)
    o.converger = int(converger[0])
    o.alpha = float(converger[1]) if (o.converger == 0 and len(converger) > 1) else 0.0
    o.use_sp2 = 1 if sp2[0] else 0
    o.sp2_eps = float(sp2[1]) if sp2[0] else 0.0
    o.max_iter = int(max_iter)
    o.warm_start = 1 if warm_start else 0
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
    plan.lib.check(
        plan.lib.dll.seqm_scf(plan.ref, C.byref(o), ptr(H), ptr(w), ptr(P), ptr(F), ptr(E), ptr(nc), ptr(ws),
                              C.byref(nit), ptr(Clast), stream_of(P)),
        "seqm_scf",
    )  # fmt: skip
    if want_C:
        return F, E, nc.to(torch.bool), int(nit.value), Clast
    return F, E, nc.to(torch.bool), int(nit.value)
