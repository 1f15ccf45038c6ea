"""scf_analytic_grad with the reference's signature (seqm/seqm_functions/anal_grad.py:16-92): gradient of the SCF
energy at a fixed density, eV/Angstrom, (nmol, molsize, 3).  One adjoint pair kernel (`seqm_gradient`) replaces
w_der / der_TETCILF / overlap_der_finiteDiff / core_core_der / contract_ao_derivatives_with_density."""
import torch

from .. import engine


def scf_analytic_grad(P0, molecule, const=None, method=None, mask=None, maskd=None, molsize=None, idxi=None,
                      idxj=None, ni=None, nj=None, xij=None, rij=None, gam=None, parnuc=None, Z=None, gss=None,
                      gpp=None, gp2=None, hsp=None, beta=None, zetas=None, zetap=None, riXH=None, ri=None):  # fmt: skip
    """Everything after `molecule` is accepted for signature compatibility and ignored: each is a field of the
    molecule's batch plan (the local-frame integrals riXH / ri are recomputed in registers)."""
    plan = molecule._plan
    xyz = molecule._refresh_geometry()
    g = engine.op_gradient(plan, xyz, engine.op_pack(plan, P0))
    grad = torch.zeros((plan.nmol * plan.molsize, 3), dtype=torch.float64, device=plan.device)
    grad[plan.real_atoms] = g
    return grad.reshape(plan.nmol, plan.molsize, 3)
