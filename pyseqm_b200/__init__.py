"""pyseqm_b200 -- B200-native (sm_100a) batched ground-state SCF path of PYSEQM behind the reference's
`Molecule` / `Electronic_Structure(seqm_parameters).forward` API.  `import pyseqm_b200 as seqm`."""
from .ElectronicStructure import Electronic_Structure  # noqa: F401
from .Molecule import Molecule  # noqa: F401
from .MolecularDynamics import KSA_XL_BOMD, XL_BOMD, Molecular_Dynamics_Basic  # noqa: F401
from .seqm_functions.constants import Constants  # noqa: F401
from .seqm_functions.read_xyz import read_xyz  # noqa: F401

__version__ = "0.1.0"
