"""MD drivers over the B200 electronic-structure step: velocity-Verlet BOMD and XL-BOMD.

Mirrors the integrator arithmetic of seqm/MolecularDynamics.py (Molecular_Dynamics_Basic.one_step 813-842,
XL_BOMD.__init__ 1330-1371, _propagate_P 1418-1427, one_step 1437-1515, initialize 1530-1605, run 958-1090).
The integrator state is a handful of AXPYs on (nmol, molsize, 3) tensors and stays in PyTorch; the XL-BOMD field
density and its history live in the packed device layout and never round-trip through the padded dense
layout inside the step loop.  HDF5/XYZ writers, checkpointing, thermostats and excited-state dynamics of the
reference are outside the accelerated path and are not provided here.
"""
import time

import torch

from . import engine
from .ElectronicStructure import Electronic_Structure

ACC_SCALE = 0.009648532800137615  # eV/A/(g/mol) -> A/fs^2       (MolecularDynamics.py:27)
VEL_SCALE = 0.9118367323190634e-3  # sqrt(K/amu) -> A/fs          (:28)
KINETIC_ENERGY_SCALE = 1.0364270099032438e2  # amu (A/fs)^2 -> eV (:29)
TEMPERATURE_SCALE = 1.160451812e4  # K/eV                          (:30)


class Molecular_Dynamics_Basic(torch.nn.Module):
    """NVE velocity Verlet with an SCF per step restarted from the previous density."""

    def __init__(self, seqm_parameters, timestep=1.0, Temp=0.0, step_offset=0, output=None, *args, **kwargs):
        super().__init__()
        self.seqm_parameters = seqm_parameters
        self.timestep = timestep
        self.Temp = Temp
        self.step_offset = step_offset
        self.esdriver = Electronic_Structure(seqm_parameters)
        self.n_dof = None
        self.history = {"Etot": [], "Ek": [], "T": []}

    def set_dof(self, molecule, constraints=0.0):
        self.n_dof = 3.0 * molecule.num_atoms - constraints

    def _kinetic_energy(self, molecule):
        return torch.sum(0.5 * molecule.mass * molecule.velocities**2, dim=(1, 2)) * KINETIC_ENERGY_SCALE

    def _calc_temperature(self, Ek):
        return Ek * TEMPERATURE_SCALE / (0.5 * self.n_dof)

    def _thermo_potential(self, molecule):
        return molecule.Etot

    def initialize_velocity(self, molecule):
        """Maxwell-Boltzmann sample rescaled to the exact temperature, centre-of-mass momentum removed
        (MolecularDynamics.py:717-744)."""
        if torch.is_tensor(molecule.velocities):
            return molecule.velocities
        if self.Temp == 0.0:
            molecule.velocities = torch.zeros_like(molecule.coordinates.detach())
            return molecule.velocities
        scale = torch.sqrt(self.Temp * molecule.mass_inverse) * VEL_SCALE
        molecule.velocities = torch.randn_like(molecule.coordinates.detach()) * scale
        T1 = self._calc_temperature(self._kinetic_energy(molecule))
        molecule.velocities.mul_(torch.sqrt(self.Temp / T1).reshape(-1, 1, 1))
        mass = molecule.mass
        M = mass.sum(dim=1, keepdim=True)
        with torch.no_grad():
            Ek0 = self._kinetic_energy(molecule)
            molecule.velocities.sub_(torch.sum(mass * molecule.velocities, dim=1, keepdim=True) / M)
            molecule.velocities.mul_((molecule.species > 0).unsqueeze(-1))
            Ek1 = self._kinetic_energy(molecule)
            molecule.velocities.mul_(torch.sqrt(Ek0 / Ek1).reshape(-1, 1, 1))
        return molecule.velocities

    def initialize(self, molecule, remove_com=None, learned_parameters=dict(), steps=None, *args, **kwargs):
        molecule.verbose = False
        if remove_com is not None:
            raise NotImplementedError("centre-of-mass removal during the run is not provided by the B200 MD driver")
        self.set_dof(molecule)
        self.initialize_velocity(molecule)
        if not torch.is_tensor(molecule.force):
            self.esdriver(molecule, learned_parameters=learned_parameters, P0=molecule.dm)
        with torch.no_grad():
            molecule.acc = molecule.force * molecule.mass_inverse * ACC_SCALE

    def one_step(self, molecule, learned_parameters=dict(), *args, **kwargs):
        dt = self.timestep
        with torch.no_grad():
            molecule.velocities.add_(0.5 * molecule.acc * dt)
            molecule.coordinates.add_(molecule.velocities * dt)
        self.esdriver(molecule, learned_parameters=learned_parameters, P0=molecule.dm, dm_prop="SCF")
        with torch.no_grad():
            molecule.acc = molecule.force * molecule.mass_inverse * ACC_SCALE
            molecule.velocities.add_(0.5 * molecule.acc * dt)

    def _do_integrator_step(self, i, molecule, learned_parameters, **kwargs):
        return self.one_step(molecule, learned_parameters=learned_parameters, **kwargs)

    def run(self, molecule, steps, learned_parameters=dict(), remove_com=None, seed=None, record=True, *args, **kwargs):
        if seed is not None:
            torch.manual_seed(int(seed))
            torch.cuda.manual_seed_all(int(seed))
        self.initialize(molecule, remove_com=remove_com, learned_parameters=learned_parameters, steps=steps)
        for i in range(self.step_offset, steps):
            self._do_integrator_step(i, molecule, learned_parameters)
            if record:
                with torch.no_grad():
                    Ek = self._kinetic_energy(molecule)
                    self.history["Ek"].append(Ek)
                    self.history["T"].append(self._calc_temperature(Ek))
                    self.history["Etot"].append(self._thermo_potential(molecule).clone())
        return molecule.coordinates, molecule.velocities, molecule.acc


class XL_BOMD(Molecular_Dynamics_Basic):
    """Extended-Lagrangian Born-Oppenheimer MD (Niklasson et al., JCP 130, 214109): no SCF inside the loop, one
    Fock build + one density solve per step, the field density P propagated with dissipation order k."""

    COEFFS = {
        3: [1.69, 150e-3, -2.0, 3.0, 0.0, -1.0],
        4: [1.75, 57e-3, -3.0, 6.0, -2.0, -2.0, 1.0],
        5: [1.82, 18e-3, -6.0, 14.0, -8.0, -3.0, 4.0, -1.0],
        6: [1.84, 5.5e-3, -14.0, 36.0, -27.0, -2.0, 12.0, -6.0, 1.0],
        7: [1.86, 1.6e-3, -36.0, 99.0, -88.0, 11.0, 32.0, -25.0, 8.0, -1.0],
        8: [1.88, 0.44e-3, -99.0, 286.0, -286.0, 78.0, 78.0, -90.0, 42.0, -10.0, 1.0],
        9: [1.89, 0.12e-3, -286.0, 858.0, -936.0, 364.0, 168.0, -300.0, 184.0, -63.0, 12.0, -1.0],
    }

    def __init__(self, damp=None, xl_bomd_params=dict(), *args, **kwargs):
        if damp is not None:
            raise NotImplementedError("Langevin damping is not provided by the B200 MD driver (damp=None only)")
        if "max_rank" in xl_bomd_params and not self._KSA:
            raise NotImplementedError("xl_bomd_params['max_rank'] selects the Krylov kernel: use KSA_XL_BOMD")
        super().__init__(*args, **kwargs)
        self.k = xl_bomd_params["k"]
        self.xl_bomd_params = xl_bomd_params
        self.m = self.k + 1
        self.kappa = self.COEFFS[self.k][0]
        self.alpha = self.COEFFS[self.k][1]
        tmp = torch.tensor(self.COEFFS[self.k][2:], dtype=torch.float64) * self.alpha
        self.coeff_D = 1.0 * self.kappa
        tmp[0] += 2.0 - self.coeff_D
        tmp[1] -= 1.0
        self.coeff = tmp.repeat(2)
        self._ctx = None

    _KSA = False

    def _thermo_potential(self, molecule):
        return molecule.Etot + molecule.Electronic_entropy

    def _propagation_source(self, ctx):
        """(X, c) of P(n+1) = kappa [c X + (1 - c) P(n)] + sum_j coeff_j Pt[j]: the density D(n) with c = 0.95
        (MolecularDynamics.py:1418-1427)."""
        return ctx["D"], 0.95

    def initialize(self, molecule, remove_com=None, learned_parameters=dict(), steps=None, *args, **kwargs):
        molecule.Electronic_entropy = torch.zeros(molecule.species.shape[0], dtype=torch.float64,
                                                  device=molecule.coordinates.device)  # fmt: skip
        if self.step_offset > 0 and self._ctx is not None and self._ctx.get("plan") is molecule._plan:
            # resume (MolecularDynamics.py:1530-1541): restored XL state (P, Pt, D) survives; only the MD bookkeeping
            # of the base class is redone
            self.set_dof(molecule)
            self.initialize_velocity(molecule)
            return
        super().initialize(molecule, remove_com=remove_com, learned_parameters=learned_parameters, steps=steps)
        plan = molecule._plan
        molecule.__dict__.pop("_C_xl", None)  # a new trajectory starts its density solves cold
        with torch.no_grad():
            dm = molecule.dm
            if molecule.orbital_stride != 4:  # method="PM6": back to the 4-slot layout the packed kernels use
                from .Molecule import narrow_orbitals

                dm = narrow_orbitals(dm, plan.molsize, molecule.orbital_stride)
            Dp = engine.op_pack(plan, dm)  # converged SCF density at t = 0
            self._ctx = {"P": Dp.clone(), "Pt": Dp.unsqueeze(0).repeat(self.m, 1), "D": Dp, "plan": plan}
        self.coeff = self.coeff.to(molecule.coordinates.device)

    def one_step(self, molecule, step, learned_parameters=dict(), *args, **kwargs):
        dt = self.timestep
        ctx = self._ctx
        plan = molecule._plan
        t0 = time.time()
        with torch.no_grad():
            molecule.velocities.add_(0.5 * molecule.acc * dt)
            molecule.coordinates.add_(molecule.velocities * dt)
            # P(n+1) = kappa [c D(n) + (1-c) P(n)] + sum_j coeff_j Pt[j]    (c = 0.95; eq. 22 of the paper)
            cindx = step % self.m
            src, c = self._propagation_source(ctx)
            P = engine.op_xl_propagate(plan, self.coeff_D, c, src, ctx["P"], ctx["Pt"],
                                       self.coeff[cindx : cindx + self.m].contiguous(), self.m - 1 - cindx)
            ctx["P"] = P
        # MD needs D, E, forces only.  molecule.dm / molecule.q are refreshed at the end of run() (or on demand by
        # refresh_density()); inside the loop the density lives in the packed layout (molecule._dm_packed)
        r = self.esdriver.conservative_force_xl.forward_packed(molecule, P, want_e=False,
                                                               learned_parameters=learned_parameters,
                                                               xl_bomd_params=self.xl_bomd_params if self._KSA else None)
        ctx["D"] = r["D"]
        if self._KSA:
            ctx["d2"] = r["dP2dt2"]
            molecule.Electronic_entropy, molecule.Krylov_Error, molecule.Fermi_occ = r["EEnt"], r["Krylov_Error"], r["Fermi_occ"]
        molecule.force, molecule.Hf, molecule.Etot = r["force"], r["Hf"], r["Etot"]
        molecule.Eelec, molecule.Enuc, molecule.Eiso = r["Eelec"], r["Enuc"], r["Eiso"]
        molecule._dm_packed = r["D"]
        with torch.no_grad():
            molecule.acc = molecule.force * molecule.mass_inverse * ACC_SCALE
            molecule.velocities.add_(0.5 * molecule.acc * dt)
        if molecule.const.do_timing:
            if molecule.coordinates.is_cuda:
                torch.cuda.synchronize()
            molecule.const.timing["MD"].append(time.time() - t0)

    def _do_integrator_step(self, i, molecule, learned_parameters, **kwargs):
        return self.one_step(molecule, i, learned_parameters=learned_parameters, **kwargs)

    def refresh_density(self, molecule):
        """Dense `molecule.dm` (and Mulliken `molecule.q`) of the latest step.  The step loop keeps the density packed;
        observers that read `molecule.dm` mid-run call this first (run() does at its end)."""
        if self._ctx is None:
            return
        molecule.dm = engine.op_unpack(molecule._plan, self._ctx["D"])
        if molecule.orbital_stride != 4:
            from .Molecule import widen_orbitals

            molecule.dm = widen_orbitals(molecule.dm, molecule._plan.molsize, molecule.orbital_stride)
        molecule.q = molecule.const.tore[molecule.species] - Electronic_Structure.atomic_charges(
            molecule.dm, n_orbital=molecule.orbital_stride)

    def run(self, molecule, steps, *args, **kwargs):
        out = super().run(molecule, steps, *args, **kwargs)
        self.refresh_density(molecule)
        return out


class KSA_XL_BOMD(XL_BOMD):
    """Krylov-subspace-approximation XL-BOMD (MolecularDynamics.py:1608-1619): the field density is driven by the rank-m
    approximation of the kernel acting on D - P at electronic temperature T_el,
    P(n+1) = kappa (dP2dt2(n) + P(n)) + sum_j coeff_j Pt[j]; xl_bomd_params = {"k", "max_rank", "err_threshold", "T_el"}."""

    _KSA = True

    def __init__(self, damp=None, xl_bomd_params=dict(), *args, **kwargs):
        for key in ("max_rank", "err_threshold", "T_el"):
            if key not in xl_bomd_params:
                raise KeyError(f"KSA_XL_BOMD needs xl_bomd_params[{key!r}]")
        super().__init__(damp, xl_bomd_params, *args, **kwargs)

    def _propagation_source(self, ctx):
        return ctx["d2"] + ctx["P"], 1.0

    def initialize(self, molecule, *args, **kwargs):
        super().initialize(molecule, *args, **kwargs)
        if "d2" not in self._ctx:
            self._ctx["d2"] = torch.zeros_like(self._ctx["P"])  # dP2dt2 = 0 at t = 0 (MolecularDynamics.py:1580-1581)

    def refresh_density(self, molecule):
        super().refresh_density(molecule)
        if self._ctx is not None:
            molecule.dP2dt2 = engine.op_unpack(molecule._plan, self._ctx["d2"])
