"""`Electronic_Structure(seqm_parameters).forward(molecule, ...)` -- the drop-in entry point
(seqm/ElectronicStructure.py:10-127).  Results are written onto `molecule` exactly as the reference does:
force, dm, Hf, Etot, Eelec, Enuc, Eiso, e_mo, e_gap, q (+ n_scf_iter); self.charge / self.notconverged."""
import torch

from .basics import Force, ForceXL
from .Molecule import reject_unsupported


class Electronic_Structure(torch.nn.Module):
    def __init__(self, seqm_parameters, *args, **kwargs):
        super().__init__()
        reject_unsupported(seqm_parameters)
        self.seqm_parameters = seqm_parameters
        self.conservative_force = Force(seqm_parameters)
        self.conservative_force_xl = ForceXL(seqm_parameters)
        self.charge = None
        self.notconverged = None

    @staticmethod
    def atomic_charges(P, n_orbital=4):
        n_molecule = P.shape[0]
        n_atom = P.shape[1] // n_orbital
        return P.diagonal(dim1=1, dim2=2).reshape(n_molecule, n_atom, n_orbital).sum(axis=2)

    def forward(self, molecule, learned_parameters=dict(), xl_bomd_params=dict(), P0=None, err_threshold=None,
                max_rank=None, T_el=None, dm_prop="SCF", *args, **kwargs):  # fmt: skip
        if max_rank is not None or T_el is not None:
            raise NotImplementedError("KSA / finite-temperature options are not part of the B200 SCF path")
        kwargs.pop("cis_amp", None)
        if dm_prop == "SCF":
            (molecule.force, P, molecule.Hf, molecule.Etot, molecule.Eelec, molecule.Enuc, molecule.Eiso, molecule.e_mo,
             molecule.e_gap, self.charge, self.notconverged) = self.conservative_force(
                molecule, P0=P0, learned_parameters=learned_parameters, *args, **kwargs)  # fmt: skip
            molecule.dm = P.detach()
        elif dm_prop == "XL-BOMD":
            (molecule.force, molecule.dm, molecule.Hf, molecule.Etot, molecule.Eelec, molecule.Enuc, molecule.Eiso,
             molecule.e_mo, molecule.e_gap, molecule.Electronic_entropy, molecule.dP2dt2, molecule.Krylov_Error,
             molecule.Fermi_occ) = self.conservative_force_xl(
                molecule, P0, learned_parameters=learned_parameters, xl_bomd_params=xl_bomd_params)  # fmt: skip
        else:
            raise NotImplementedError(f"dm_prop={dm_prop!r} is not implemented by the B200 path")
        with torch.no_grad():
            molecule.q = molecule.const.tore[molecule.species] - self.atomic_charges(molecule.dm)

    def get_force(self):
        return self.force
