"""`Molecule`: the batch container with the reference's attribute contract (seqm/Molecule.py:11-206).

Construction runs the parser (pair list and index maps, seqm/basics.py:219-403) and the parameter gather
(seqm/basics.py:406-448) as torch index arithmetic on the molecule's device and builds the device batch
plan (`_plan`) that every kernel call uses.  Results are delivered by mutation, as in the reference.
"""
from typing import Optional

import torch

from . import engine
from ._lib import get_lib

SUPPORTED_METHODS = ("MNDO", "AM1", "PM3", "PM6_SP", "PM6")


def pm6_d_shell(species):
    """Elements that carry a d shell in the reference's PM6 (the nSuperHeavy set, seqm/basics.py:258-269)."""
    s = species
    return (((s > 12) & (s < 18)) | ((s > 20) & (s < 30)) | ((s > 32) & (s < 36)) | ((s > 38) & (s < 48))
            | ((s > 50) & (s < 54)) | ((s > 70) & (s < 80)) | (s == 57))  # fmt: skip


def widen_orbitals(P4, molsize, stride=9):
    """(nmol, 4 molsize, 4 molsize) -> (nmol, 9 molsize, 9 molsize): the reference's PM6 layout keeps 9 orbital
    slots per atom (s, px, py, pz, 5 d) and leaves the d slots of sp-only atoms zero (packd.py:195-218)."""
    b = P4.shape[0]
    out = torch.zeros((b, molsize, stride, molsize, stride), dtype=P4.dtype, device=P4.device)
    out[:, :, :4, :, :4] = P4.view(b, molsize, 4, molsize, 4)
    return out.view(b, stride * molsize, stride * molsize)


def narrow_orbitals(P9, molsize, stride=9):
    b = P9.shape[0]
    return P9.reshape(b, molsize, stride, molsize, stride)[:, :, :4, :, :4].reshape(b, 4 * molsize, 4 * molsize).contiguous()


check_input = engine.check_input  # rows must be non-increasing in Z (Molecule.py:188-206); run by the batch plan


def reject_unsupported(seqm_parameters):
    """The fast path never falls back silently (SURVEY 8(b) option matrix)."""
    sp = seqm_parameters
    if sp.get("method") not in SUPPORTED_METHODS:
        raise NotImplementedError(f"method {sp.get('method')!r}: the B200 path implements {SUPPORTED_METHODS}")
    bad = []
    if sp.get("UHF", False):
        bad.append("UHF=True")
    if sp.get("excited_states"):
        bad.append("excited_states")
    if sp.get("active_state", 0):
        bad.append("active_state>0")
    if sp.get("scf_backward", 0) not in (0,):
        bad.append("scf_backward in {1,2}")
    if sp.get("2nd_grad", False):
        bad.append("2nd_grad")
    if sp.get("dispersion", False):
        bad.append("dispersion")
    if sp.get("normal modes", False):
        bad.append("normal modes")
    conv = sp.get("scf_converger", [2])
    if conv[0] not in (0, 1, 2, 3):
        bad.append(f"scf_converger={conv}")
    if conv[0] == 3 and not (len(conv) > 1 and isinstance(conv[1], dict) and {"max_rank", "err_threshold", "T_el"} <= set(conv[1])):
        bad.append("scf_converger=[3] without its {'max_rank', 'err_threshold', 'T_el'} dictionary")
    if bad:
        raise NotImplementedError("not implemented by the B200 SCF path: " + ", ".join(bad))


class Molecule(torch.nn.Module):
    def __init__(self, const, seqm_parameters, coordinates, species, charges=0, mult=1, learned_parameters=dict(),
                 do_large_tensors=True, _lib=None, *args, **kwargs):  # fmt: skip
        super().__init__()
        self.const = const
        reject_unsupported(seqm_parameters)
        if not species.is_cuda:
            check_input(species)  # on the GPU the batch plan folds this test into its single host read-back
        if coordinates.dtype != torch.float64:
            raise NotImplementedError("the B200 path is fp64 only: pass float64 coordinates")
        self.species = species
        self.coordinates = torch.nn.Parameter(coordinates)
        self.coordinates.requires_grad_(False)
        # tot_charge / mult and most index / mass attributes below are materialised on first access (__getattr__): the SCF
        # path never reads them, and every eager tensor op here is host time inside an end-to-end step
        self.__dict__["_charges_in"], self.__dict__["_mult_in"] = charges, mult
        self.seqm_parameters = seqm_parameters
        self.method = seqm_parameters["method"]
        if callable(learned_parameters):
            raise NotImplementedError("callable learned_parameters need autograd through the SCF; not on the B200 path")
        # method="PM6": a batch with d-shell elements (the reference's nSuperHeavy set) runs the spd kernels with 9 orbitals on
        # those atoms; a batch without any is numerically PM6_SP with the PM6 parameter file, its results widened to the
        # reference's 9-slot layout.
        self.orbital_stride = 9 if self.method == "PM6" else 4
        kernel_method, table = self.method, None
        if self.method == "PM6":
            kernel_method, table = ("PM6_D", "PM6") if bool(pm6_d_shell(species).any()) else ("PM6_SP", "PM6")
            if kernel_method == "PM6_D" and seqm_parameters.get("learned"):
                raise NotImplementedError("learned parameters with PM6 d-shell elements are not on the B200 path")
        lib = _lib if _lib is not None else get_lib()
        # basics.py:442-448: only the names listed in seqm_parameters["learned"] are taken from learned_parameters,
        # everything else comes from the method's table.  Values only: no gradients flow back to them.
        learned = {}
        for name in seqm_parameters.get("learned", []):
            t = learned_parameters[name]
            if t.requires_grad:
                raise NotImplementedError("gradients with respect to learned parameters need autograd through the SCF; not on the B200 path")
            learned[name] = t.detach()
        plan = engine.BatchPlan(lib, species, kernel_method, parameters=learned, charges=charges if torch.is_tensor(charges) else int(charges), table=table,
                                outer_cutoff=seqm_parameters.get("pair_outer_cutoff", 1.0e10))
        self._plan = plan
        if seqm_parameters.get("elements") is None:
            seqm_parameters["elements"] = plan.elements
        dev = coordinates.device
        self.nmol, self.molsize = plan.nmol, plan.molsize
        # nHeavy / nHydro / nocc / nSuperHeavy / Z / atom_molid / idxi / idxj / norb / num_atoms / mass / mass_inverse and
        # maskd / mask / mask_l / ni / nj / pair_molid / xij / rij are derived on first access (see __getattr__):
        # the kernels never read them, they exist for the reference's attribute contract
        # pair_outer_cutoff (basics.py:209, 326): the pair list stays the dense triangular one; a pair at or beyond the
        # cutoff contributes exactly nothing (the kernels return zero for its w, overlap block, core-core energy and
        # gradient), which is what dropping it from the reference's list does.  idxi/idxj/rij/xij/w keep such pairs.
        # per-atom parameter dict (Molecule.py:86-115)
        names = ["U_ss", "U_pp", "zeta_s", "zeta_p", "beta_s", "beta_p", "g_ss", "g_sp", "g_pp", "g_p2", "h_sp", "alpha"]
        ng = {"MNDO": 0, "AM1": 4, "PM3": 2, "PM6_SP": 4, "PM6": 4}[self.method]
        for g in range(1, ng + 1):
            names += [f"Gaussian{g}_K", f"Gaussian{g}_L", f"Gaussian{g}_M"]
        self.parameters = {k: plan.parameter(k) for k in names}
        self.parameters["beta"] = torch.stack((self.parameters["beta_s"], self.parameters["beta_p"]), dim=1)
        zeros = torch.zeros_like(self.parameters["zeta_s"])
        for k in ("zeta_d", "s_orb_exp_tail", "p_orb_exp_tail", "d_orb_exp_tail", "U_dd", "F0SD", "G2SD", "rho_core"):
            self.parameters[k] = zeros
        if plan.d_mode:
            for k in ("zeta_d", "U_dd", "beta_d"):
                self.parameters[k] = plan.parameter(k)
            self.parameters["beta"] = torch.stack((self.parameters["beta_s"], self.parameters["beta_p"],
                                                   self.parameters["beta_d"]), dim=1)  # fmt: skip
        self.parameters["Kbeta"] = None
        zmax = plan.zmax
        if plan.pw is not None:
            self.alp, self.chi = plan.pw[0], plan.pw[1]
        else:
            self.alp = torch.zeros((zmax + 1, zmax + 1), dtype=torch.float64, device=dev)
            self.chi = torch.zeros_like(self.alp)

        self.force = None
        self.velocities = None
        self.acc = None
        self.dm: Optional[torch.Tensor] = None
        self.q: Optional[torch.Tensor] = None
        self.w: Optional[torch.Tensor] = None
        self._parnuc = None
        self._gam = None
        self.Hf = self.Etot = self.Eelec = self.Enuc = self.Eiso = None
        self.e_mo = self.e_gap = None
        self.molecular_orbitals = None
        self.charge = None
        self.dipole = None
        self.verbose = True
        self.analytical_gradient = None
        self.active_state = 0
        self.Electronic_entropy = self.Fermi_occ = self.dP2dt2 = self.Krylov_Error = None
        self.cis_amplitudes = None
        self.n_scf_iter: Optional[int] = None  # the count the reference only prints (scf_loop.py:975-992)

    _LAZY = ("maskd", "mask", "mask_l", "ni", "nj", "pair_molid", "xij", "rij", "w")
    _LAZY2 = ("tot_charge", "mult", "nHeavy", "nHydro", "nocc", "nSuperHeavy", "Z", "atom_molid", "idxi", "idxj", "norb",
              "num_atoms", "mass", "mass_inverse")  # fmt: skip

    def _lazy2(self, name):
        d = self.__dict__
        plan, dev = d["_plan"], self.coordinates.device
        if name in ("tot_charge", "mult"):
            v = d["_charges_in" if name == "tot_charge" else "_mult_in"]
            d[name] = v if torch.is_tensor(v) else v * torch.ones(self.coordinates.shape[0], device=dev)
        elif name in ("nHeavy", "nHydro", "nocc", "nSuperHeavy", "norb"):
            d["nHydro"], d["nocc"] = plan.nhyd, plan.nocc
            if plan.d_mode:  # basics.py:239-269: nHeavy counts the sp-only heavy atoms, nSuperHeavy the d-shell ones
                d["nSuperHeavy"], d["nHeavy"] = plan.nsh, plan.nheavy - plan.nsh
            else:
                d["nSuperHeavy"], d["nHeavy"] = torch.zeros_like(plan.nheavy), plan.nheavy
            d["norb"] = d["nHydro"] + 4 * d["nHeavy"] + 9 * d["nSuperHeavy"]
        elif name in ("Z", "atom_molid", "idxi", "idxj"):
            d["Z"], d["atom_molid"], d["idxi"], d["idxj"] = plan.Z, plan.atom_mol, plan.pair_i, plan.pair_j
        else:
            non_zero = self.species != 0
            d["num_atoms"] = non_zero.sum(dim=1).to(self.coordinates.dtype)
            d["mass"] = self.const.mass[self.species].unsqueeze(2)
            d["mass_inverse"] = torch.where(non_zero.unsqueeze(2), 1.0 / d["mass"].clamp_min(1e-300), torch.zeros_like(d["mass"]))
        return d[name]

    def __getattr__(self, name):
        if name in Molecule._LAZY2 and "_plan" in self.__dict__:
            return self._lazy2(name)
        if name in Molecule._LAZY and "_plan" in self.__dict__:
            plan = self.__dict__["_plan"]
            ms, pos = plan.molsize, plan.atom_local
            d = self.__dict__
            if name == "w":  # method="PM6": the (npairs, 45, 45) tensor of the reference is assembled on first access
                parts = d.get("_w_parts")
                if parts is None:
                    raise AttributeError(name)
                d["w"] = engine.dense_w45(plan, *parts)
                return d["w"]
            if name in ("xij", "rij"):  # unit vector i->j and distance in bohr (basics.py:737-746)
                xyz = plan.real_xyz(self.coordinates)
                dv = xyz[plan.pair_j] - xyz[plan.pair_i]
                dist = torch.linalg.norm(dv, dim=1)
                d["xij"] = dv / dist.unsqueeze(1)
                d["rij"] = dist * self.const.length_conversion_factor
            elif name == "maskd":
                d["maskd"] = plan.atom_mol * ms * ms + pos * (ms + 1)
            elif name in ("ni", "nj"):
                d["ni"], d["nj"] = plan.Z[plan.pair_i], plan.Z[plan.pair_j]
            else:
                pm = plan.atom_mol[plan.pair_i]
                d["pair_molid"] = pm
                d["mask"] = pm * ms * ms + pos[plan.pair_i] * ms + pos[plan.pair_j]
                d["mask_l"] = pm * ms * ms + pos[plan.pair_j] * ms + pos[plan.pair_i]
            return d[name]
        return super().__getattr__(name)

    def _refresh_geometry(self):
        """Packed coordinates of the real atoms; the cached pair geometry (xij, rij) is invalidated."""
        self.__dict__.pop("xij", None)
        self.__dict__.pop("rij", None)
        return self._plan.real_xyz(self.coordinates)

    def get_coordinates(self):
        return self.coordinates

    def get_species(self):
        return self.species
