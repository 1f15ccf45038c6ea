"""fock(...) with the reference's positional signature (seqm/seqm_functions/fock.py:132-181)."""
import torch

from .. import engine
from ._plans import plan_from_reference_args, plan_of
from .hcore import dense_from_blocks


def fock(nmol, molsize, P0, M, maskd, mask, idxi, idxj, w, W, gss, gpp, gsp, gp2, hsp, themethod, zetas, zetap,
         zetad, Z, F0SD, G2SD):  # fmt: skip
    """F = Hcore + G(P): P0 dense (nmol, 4 molsize, 4 molsize), M the block-form Hcore of `hcore`, w (npairs, 10, 10).
    Returns the full symmetric dense F.  The batch plan travels on M / w when they come from `hcore(molecule)`;
    otherwise it is rebuilt from (maskd, Z) and the per-atom one-centre parameters passed here."""
    if themethod == "PM6":
        raise NotImplementedError("9-orbital PM6 blocks are not on the B200 path (SURVEY a17)")
    plan = plan_of(M, w)
    if plan is None:
        plan = plan_from_reference_args(nmol, molsize, maskd, Z, themethod, g_ss=gss, g_pp=gpp, g_sp=gsp, g_p2=gp2,
                                        h_sp=hsp, zeta_s=zetas, zeta_p=zetap)  # fmt: skip
    H = getattr(M, "_seqm_H", None)
    if H is None:
        Hd = dense_from_blocks(M, nmol, molsize)
        Hd = Hd.triu() + Hd.triu(1).transpose(1, 2)
        H = engine.op_pack(plan, Hd)
    F = engine.op_fock(plan, engine.op_pack(plan, P0), H, w.contiguous())
    return engine.op_unpack(plan, F)
