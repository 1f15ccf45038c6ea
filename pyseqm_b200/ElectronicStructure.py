"""`Electronic_Structure(seqm_parameters).forward(molecule, ...)` -- the drop-in entry point
(seqm/ElectronicStructure.py:10-127).  Results are written onto `molecule` exactly as the reference does:
force, dm, Hf, Etot, Eelec, Enuc, Eiso, e_mo, e_gap, q (+ n_scf_iter); self.charge / self.notconverged."""
import torch

from .basics import Force, ForceXL
from .Molecule import reject_unsupported


def orbital_charge_table(v, nHeavy, nHydro, molsize):
    """Literal restatement of scf_loop.py:2346-2387 (closed shell): with v2 = v**2,
       charge[i, :norb, :nH]      = v2[i][:norb, :4 nH].reshape(norb, 4, nH).sum(1)
       charge[i, :norb, nH:nH+ny] = v2[i][:norb, 4 nH : 4 nH + ny]
    grouped by (nHeavy, nHydro) so that each group is one batched tensor expression."""
    nmol = v.shape[0]
    out = torch.zeros((nmol, 4 * molsize, molsize), dtype=v.dtype, device=v.device)
    key = nHeavy * 1000 + nHydro
    for k in torch.unique(key).tolist():
        nh, ny = k // 1000, k % 1000
        sel = torch.nonzero(key == k, as_tuple=False).squeeze(1)
        norb = 4 * nh + ny
        v2 = v[sel][:, :norb, :norb] ** 2
        if nh > 0:
            out[sel, :norb, :nh] = v2[:, :, : 4 * nh].reshape(-1, norb, 4, nh).sum(dim=2)
        if ny > 0:
            out[sel, :norb, nh : nh + ny] = v2[:, :, 4 * nh : 4 * nh + ny]
    return out


class Electronic_Structure(torch.nn.Module):
    def __init__(self, seqm_parameters, *args, **kwargs):
        super().__init__()
        reject_unsupported(seqm_parameters)
        self.seqm_parameters = seqm_parameters
        self.conservative_force = Force(seqm_parameters)
        self.conservative_force_xl = ForceXL(seqm_parameters)
        self._charge = None
        self._charge_src = None
        self.notconverged = None

    @property
    def charge(self):
        """Per-orbital atomic "charge" table (nmol, 4*molsize, molsize) of scf_loop.py:2346-2387, evaluated lazily
        from the eigenvectors of the last forward (it is a by-product nobody on the hot path consumes)."""
        if self._charge is None and self._charge_src is not None:
            mol, mo = self._charge_src  # index tensors of the molecule are themselves built on first access
            self._charge = orbital_charge_table(mo, mol.nHeavy, mol.nHydro, mol.molsize)
        return self._charge

    @charge.setter
    def charge(self, value):
        self._charge = value

    @staticmethod
    def atomic_charges(P, n_orbital=4):
        n_molecule = P.shape[0]
        n_atom = P.shape[1] // n_orbital
        return P.diagonal(dim1=1, dim2=2).reshape(n_molecule, n_atom, n_orbital).sum(axis=2)

    def forward(self, molecule, learned_parameters=dict(), xl_bomd_params=dict(), P0=None, err_threshold=None,
                max_rank=None, T_el=None, dm_prop="SCF", *args, **kwargs):  # fmt: skip
        # max_rank / T_el: accepted and unused, as in the reference (ElectronicStructure.py:47-48); the KSA parameters travel
        # in xl_bomd_params (XL-BOMD) or in scf_converger = [3, {...}] (SCF)
        kwargs.pop("cis_amp", None)
        if dm_prop == "SCF":
            (molecule.force, P, molecule.Hf, molecule.Etot, molecule.Eelec, molecule.Enuc, molecule.Eiso, molecule.e_mo,
             molecule.e_gap, _, self.notconverged) = self.conservative_force(
                molecule, P0=P0, learned_parameters=learned_parameters, *args, **kwargs)  # fmt: skip
            molecule.dm = P.detach()
            self._charge = None
            self._charge_src = None
            if molecule.molecular_orbitals is not None and molecule.method != "PM6":  # scf_loop.py:2350: none for PM6
                self._charge_src = (molecule, molecule.molecular_orbitals)
        elif dm_prop == "XL-BOMD":
            (molecule.force, molecule.dm, molecule.Hf, molecule.Etot, molecule.Eelec, molecule.Enuc, molecule.Eiso,
             molecule.e_mo, molecule.e_gap, molecule.Electronic_entropy, molecule.dP2dt2, molecule.Krylov_Error,
             molecule.Fermi_occ) = self.conservative_force_xl(
                molecule, P0, learned_parameters=learned_parameters, xl_bomd_params=xl_bomd_params)  # fmt: skip
        else:
            raise NotImplementedError(f"dm_prop={dm_prop!r} is not implemented by the B200 path")
        q = molecule.__dict__.pop("_q_post", None)  # written by the post-SCF kernel of the forward that just ran
        if q is not None:
            molecule.q = q
        else:
            with torch.no_grad():
                molecule.q = molecule.const.tore[molecule.species] - self.atomic_charges(
                    molecule.dm, n_orbital=getattr(molecule, "orbital_stride", 4))

    def get_force(self):
        return self.force
