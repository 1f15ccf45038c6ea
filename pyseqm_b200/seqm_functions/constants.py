"""`Constants` with the reference's attribute contract (seqm/seqm_functions/constants.py:26-234).
The tables are data extracted from the reference by tools/make_tables.py (pyseqm_b200/data/element_tables.json)."""
import torch

from ..engine import element_tables

_el = element_tables()
ev = _el["ev"]  # 27.21, constants.py:4
a0 = _el["a0"]  # 0.529167, constants.py:9
ev_kcalpmol = _el["ev_kcalpmol"]
overlap_cutoff = _el["overlap_cutoff"]
to_debye = _el["to_debye"]
debye_to_AU = _el["debye_to_AU"]


class Constants(torch.nn.Module):
    def __init__(self, do_timing=False, length_conversion_factor=(1.0 / a0), energy_conversion_factor=1.0):
        super().__init__()
        self.length_conversion_factor = length_conversion_factor
        self.energy_conversion_factor = energy_conversion_factor
        self.label = list(_el["label"])
        f64 = lambda k: torch.nn.Parameter(torch.tensor(_el[k], dtype=torch.float64), requires_grad=False)  # noqa: E731
        for k in ("atomic_num", "tore", "iso", "qn", "ussc", "uppc", "gssc", "gspc", "hspc", "gp2c", "gppc", "eheat", "mass"):
            setattr(self, k, f64(k))
        self.qn_int = torch.nn.Parameter(torch.tensor(_el["qn_int"], dtype=torch.int64), requires_grad=False)
        self.qnD_int = torch.nn.Parameter(torch.tensor(_el["qnD_int"], dtype=torch.int64), requires_grad=False)
        self.do_timing = do_timing
        if do_timing:
            self.timing = {"Hcore + STO Integrals": [], "SCF": [], "Force": [], "MD": [], "D*": [], "CIS/RPA": []}

    def forward(self):
        pass
