"""scf_loop(molecule, ...) with the reference's signature and 12-tuple (seqm/seqm_functions/scf_loop.py:2034-2394)."""
import torch

from .. import engine
from ..ElectronicStructure import orbital_charge_table
from .hcore import hcore


def scf_loop(molecule, eps=1.0e-4, P=None, sp2=[False], scf_converger=[1], eig=False, scf_backward=0,
             scf_backward_eps=1.0e-2):  # fmt: skip
    """Returns F, e, P, Hcore, w, charge, rho0xi, rho0xj, riXH, ri, notconverged, v (dense layouts; Hcore in the
    reference's block form; e / charge / v are None unless `eig`).  A given P is updated in place."""
    if scf_backward not in (0,):
        raise NotImplementedError("scf_backward in {1, 2} needs autograd through the SCF; not on the B200 path")
    plan = molecule._plan
    M, w, rho0xi, rho0xj, riXH, ri = hcore(molecule)
    H = M._seqm_H
    Pp = engine.op_initial_density(plan) if P is None else engine.op_pack(plan, P)
    F, Eelec, notconverged, n_iter, Clast = engine.op_scf(plan, H, w, Pp, eps, scf_converger, sp2, want_C=True)
    molecule.n_scf_iter = n_iter
    Pd = engine.op_unpack(plan, Pp, out=P if (P is not None and P.is_contiguous()) else None)
    Fd = engine.op_unpack(plan, F)
    e = charge = v = None
    if eig:
        e_n, _, Cm = engine.op_eig_density(plan, F, want_P=False, want_C=True, Cguess=Clast)
        e = torch.zeros((plan.nmol, 4 * plan.molsize), dtype=torch.float64, device=plan.device)
        e[:, : plan.nmax] = e_n
        v = engine.op_orbitals_dense(plan, Cm)
        charge = orbital_charge_table(v, plan.nheavy, plan.nhyd, plan.molsize)
    return Fd, e, Pd, M, w, charge, rho0xi, rho0xj, riXH, ri, notconverged, v
