"""Batch plans for the operator-level (level B) entry points, which receive the reference's positional
tensors instead of a `Molecule`.  A plan is recovered from, in this order: the `_seqm_plan` tag that
`hcore(molecule)` leaves on its outputs, or a rebuild from the index tensors the reference passes."""
import torch

from .. import engine
from .._lib import get_lib

_MATRIX_PLANS = {}
_LIB = None


def use_library(lib):
    """Tests point the operator-level entry points at the host-emulation build; the default is the CUDA library."""
    global _LIB
    _LIB = lib
    _MATRIX_PLANS.clear()


def _library(lib=None):
    return lib or _LIB or get_lib()


def tag(t, plan):
    t._seqm_plan = plan
    return t


def plan_of(*tensors):
    for t in tensors:
        p = getattr(t, "_seqm_plan", None)
        if p is not None:
            return p
    return None


def matrix_plan(nHeavy, nHydro, nocc=None, lib=None):
    """Plan that only describes matrix shapes (for pack / unpack / sym_eig_trunc / SP2): molecules are stood in
    for by nHeavy carbon and nHydro hydrogen atoms; the occupation is taken from `nocc`."""
    dev = nHeavy.device
    nh, ny = nHeavy.to(torch.int64), nHydro.to(torch.int64)
    key = (str(dev), tuple(nh.tolist()), tuple(ny.tolist()), None if nocc is None else tuple(nocc.to(torch.int64).tolist()))
    plan = _MATRIX_PLANS.get(key)
    if plan is None:
        ms = int((nh + ny).max())
        ar = torch.arange(ms, device=dev).unsqueeze(0)
        species = torch.where(ar < nh.unsqueeze(1), 6, torch.where(ar < (nh + ny).unsqueeze(1), 1, 0)).to(torch.int64)
        nel = 4 * nh + ny
        charges = (nel % 2) if nocc is None else nel - 2 * nocc.to(torch.int64)
        plan = engine.BatchPlan(_library(lib), species, "AM1", charges=charges)
        if len(_MATRIX_PLANS) > 64:
            _MATRIX_PLANS.clear()
        _MATRIX_PLANS[key] = plan
    return plan


def dense_index(plan, size=None):
    """(nmol, nmax) position of every packed orbital in the padded dense matrix, and its validity mask."""
    nmax = plan.nmax
    k = torch.arange(nmax, device=plan.device).unsqueeze(0)
    h4 = (4 * plan.nheavy).unsqueeze(1)
    idx = torch.where(k < h4, k, h4 + 4 * (k - h4))
    valid = k < plan.norb.unsqueeze(1)
    return torch.where(valid, idx, 0), valid


def plan_from_reference_args(nmol, molsize, maskd, Z, themethod, lib=None, **per_atom):
    """Rebuild the batch plan from the reference's flat index tensors (basics.py:219-403): `maskd` is the flat
    index of each real atom's diagonal block in (nmol * molsize**2)."""
    ms2 = molsize * molsize
    mol = torch.div(maskd, ms2, rounding_mode="floor")
    pos = torch.div(maskd - mol * ms2, molsize + 1, rounding_mode="floor")
    species = torch.zeros((nmol, molsize), dtype=torch.int64, device=Z.device)
    species[mol, pos] = Z.to(torch.int64)
    pars = {k: v for k, v in per_atom.items() if v is not None}
    return engine.BatchPlan(_library(lib), species, themethod, parameters=pars)
