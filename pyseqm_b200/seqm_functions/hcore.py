"""hcore(molecule) with the reference's signature and return layout (seqm/seqm_functions/hcore.py:9-179)."""
import torch

from .. import engine
from ._plans import tag


def blocks_from_dense(Hd, nmol, molsize, nbf=4):
    return Hd.view(nmol, molsize, nbf, molsize, nbf).transpose(2, 3).reshape(-1, nbf, nbf)


def dense_from_blocks(M, nmol, molsize, nbf=4):
    n = nbf * molsize
    return M.view(nmol, molsize, molsize, nbf, nbf).transpose(2, 3).reshape(nmol, n, n)


def hcore(molecule, doTETCI=True):
    """Returns (M, w, rho0xi, rho0xj, riXH, ri): M (nmol * molsize**2, 4, 4) holds the upper triangle of Hcore in
    the reference's block form, w (npairs, 10, 10).  riXH / ri (the local-frame integrals the reference hands to its
    analytic gradient) are not materialised -- `scf_analytic_grad` recomputes them in registers -- and are None."""
    if molecule.orbital_stride != 4:
        raise NotImplementedError("operator-level hcore is sp only; method='PM6' is served through Electronic_Structure")
    plan = molecule._plan
    xyz = molecule._refresh_geometry()
    w, hab = engine.op_pair_integrals(plan, xyz)
    H = engine.op_hcore(plan, w, hab)
    Hd = engine.op_unpack(plan, H).triu()
    M = blocks_from_dense(Hd, plan.nmol, plan.molsize).contiguous()
    rc, rho0 = plan.parameter("rho_core"), plan.parameter("rho0")
    rho0 = torch.where(rc != 0.0, rc, rho0)  # two_elec_two_center_int.py:273-281
    tag(M, plan)
    tag(w, plan)
    M._seqm_H = H
    return M, w, rho0[plan.pair_i], rho0[plan.pair_j], None, None
