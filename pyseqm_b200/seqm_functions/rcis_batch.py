"""makeA_pi_batched(...) with the reference's signature (seqm/seqm_functions/rcis_batch.py:296-403): the CIS sigma-vector
building block, i.e. the contraction of AO transition densities with the two-electron integrals."""
import torch

from .. import engine


def makeA_pi_batched(mol, P_xi, w_=None, allSymmetric=False):
    """P_xi (nmol, nroots, norb, norb): per molecule and root a (generally non-symmetric) AO matrix in the molecule's packed
    orbital order [4 per heavy atom][1 per hydrogen]; like the reference, every molecule of the batch must have the same
    (nHeavy, nHydro).  Returns sum_jb (mu nu || jb) X_jb as (nmol, nroots, norb, norb).  w_: the (npairs, 10, 10) integrals of
    `hcore(mol)` (default: those of the molecule's last forward).  allSymmetric: the caller guarantees symmetric inputs."""
    plan = mol._plan
    if plan.d_mode:
        raise NotImplementedError("CIS contractions with PM6 d orbitals are not on the B200 path")
    nmol, nroots, norb = P_xi.shape[0], P_xi.shape[1], P_xi.shape[2]
    if nmol != plan.nmol or bool((plan.norb != norb).any()):
        raise ValueError("makeA_pi_batched needs a batch of molecules with identical orbital counts (rcis_batch.py:310-312)")
    w = mol.w if w_ is None else w_
    w = w.reshape(-1, 10, 10).contiguous()
    out = torch.empty((nmol, nroots, norb, norb), dtype=torch.float64, device=plan.device)
    for r in range(nroots):
        X = P_xi[:, r].contiguous().reshape(-1)
        out[:, r] = engine.op_sigma_ao(plan, X, w, all_symmetric=allSymmetric).reshape(nmol, norb, norb)
    return out
