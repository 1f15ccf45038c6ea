"""Energy terms with the reference's signatures (seqm/seqm_functions/energy.py:8-216).  The dense-matrix forms
are evaluated by the packed kernels when the inputs can be packed; the per-atom sums are index arithmetic."""
import torch

from .. import engine
from ._plans import plan_of


def elec_energy_isolated_atom(const, Z, uss, upp, gss, gpp, gsp, gp2, hsp):
    return (uss * const.ussc[Z] + upp * const.uppc[Z] + gss * const.gssc[Z] + gpp * const.gppc[Z]
            + gsp * const.gspc[Z] + gp2 * const.gp2c[Z] + hsp * const.hspc[Z])  # fmt: skip


def elec_energy(P, F, Hcore, doTriu=True, molecule=None):
    """Eelec = 1/2 sum P o (h + F), h symmetrised from the upper-triangular Hcore (energy.py:26-53).  With
    `molecule=` (or tensors tagged by hcore) the packed `seqm_elec_energy` kernel does the reduction."""
    h = Hcore.triu() + Hcore.triu(1).transpose(1, 2) if doTriu else Hcore
    plan = molecule._plan if molecule is not None else plan_of(P, F, Hcore)
    if plan is None:
        return 0.5 * torch.sum(P * (h + F), dim=(1, 2))
    return engine.op_elec_energy(plan, engine.op_pack(plan, P), engine.op_pack(plan, h), engine.op_pack(plan, F))


def elec_energy_xl(D, P, F, Hcore, molecule=None):
    h = Hcore.triu() + Hcore.triu(1).transpose(1, 2)
    plan = molecule._plan if molecule is not None else plan_of(D, P, F, Hcore)
    if plan is None:
        return torch.sum(D * F - 0.5 * (F - h) * P, dim=(1, 2))
    pk = lambda t: engine.op_pack(plan, t)  # noqa: E731
    return engine.op_elec_energy_xl(plan, pk(D), pk(P), pk(F), pk(h))


def pair_nuclear_energy(molecule, w=None):
    """Core-core repulsion per pair, (npairs,).  The reference's 15-argument form (energy.py:91-174) is reduced to
    the molecule: every one of those arguments is a field of it."""
    plan = molecule._plan
    xyz = molecule._refresh_geometry()
    if w is None:
        w, _ = engine.op_pair_integrals(plan, xyz)
    if w.shape[-1] != 10:
        w = w[:, :10, :10].transpose(1, 2).contiguous()
    EnucAB, _ = engine.op_nuclear_energy(plan, xyz, w)
    return EnucAB


def total_energy(nmol, pair_molid, EnucAB, Eelec):
    Enuc = torch.zeros((nmol,), dtype=EnucAB.dtype, device=EnucAB.device)
    Enuc.index_add_(0, pair_molid, EnucAB)
    return Eelec + Enuc, Enuc


def heat_formation(const, nmol, atom_molid, Z, Etot, Eiso, flag=True):
    Eiso_sum = torch.zeros_like(Etot).index_add_(0, atom_molid, Eiso)
    if flag:
        eheat_sum = torch.zeros_like(Etot).index_add_(0, atom_molid, const.eheat[Z])
        return Etot - Eiso_sum + eheat_sum, Eiso_sum
    return Etot - Eiso_sum, Eiso_sum
