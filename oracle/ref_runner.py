"""oracle/_ref: the UNMODIFIED reference (lanl/PYSEQM v2.0.0), installed by recipe, and a runner for it.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rules as seqm_oracle: tests/, __graft_entry__ and bench.py's
cpu_baseline / --impl reference legs; never the product).

Recipe (`install_reference`, run by `__graft_entry__.build()` in the build container where /root/reference exists):
    pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference in /tmp>
(/root/reference is read-only, setuptools writes build/ and egg-info into the source tree, hence the copy;
--no-deps because h5py is not installable offline -- `oracle/h5py_stub` stands in for it: the reference imports h5py
at module load for its MD writers only).  `oracle/_ref/` is git-ignored (no reference source enters the history) but
NOT gpurun-ignored, so it travels to the GPU box with the snapshot like the built .so files do.  Nothing at run time
reads /root/reference.
"""
import contextlib
import io
import os
import re
import shutil
import subprocess
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(HERE, "h5py_stub")
SOURCE = "/root/reference"


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "seqm", "__init__.py"))


def install_reference(force=False, verbose=False):
    """Install the reference into oracle/_ref (idempotent).  Returns the path or None when /root/reference is absent."""
    if reference_available() and not force:
        return REF_DIR
    if not os.path.isdir(os.path.join(SOURCE, "seqm")):
        return None
    tmp = tempfile.mkdtemp(prefix="pyseqm_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns("model.pt", ".git", "docs"))
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_DIR, src]  # fmt: skip
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not reference_available():
            if verbose:
                print("pip install of the reference failed; copying the package directory instead\n", r.stderr[-2000:])
            os.makedirs(REF_DIR, exist_ok=True)
            shutil.copytree(os.path.join(SOURCE, "seqm"), os.path.join(REF_DIR, "seqm"), dirs_exist_ok=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return REF_DIR if reference_available() else None


def import_reference():
    """Import `seqm` from oracle/_ref (h5py stubbed).  Raises if the install is missing."""
    if not reference_available():
        raise RuntimeError("oracle/_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` in the build "
                           "container (it installs /root/reference there by recipe)")
    for p in (STUB_DIR, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import torch

    torch.set_default_dtype(torch.float64)
    import seqm  # noqa: F401
    from seqm.ElectronicStructure import Electronic_Structure
    from seqm.Molecule import Molecule
    from seqm.seqm_functions.constants import Constants

    return Constants, Molecule, Electronic_Structure


def run_reference(species, coordinates, seqm_parameters, device="cpu", threads=None, want=("Etot", "Hf", "force", "dm")):
    """One Electronic_Structure.forward of the reference.  Returns (dict of numpy results incl. n_scf_iter, seconds of the
    forward call alone -- Molecule() construction excluded, device synchronised on both sides)."""
    import time

    import numpy as np
    import torch

    Constants, Molecule, Electronic_Structure = import_reference()
    if threads:
        torch.set_num_threads(int(threads))
    dev = torch.device(device)
    sp = dict(seqm_parameters)
    s = torch.as_tensor(np.asarray(species), dtype=torch.int64, device=dev)
    c = torch.as_tensor(np.asarray(coordinates), dtype=torch.float64, device=dev)
    const = Constants().to(dev)
    mol = Molecule(const, sp, c, s).to(dev)
    es = Electronic_Structure(sp).to(dev)
    buf = io.StringIO()
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        es(mol)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    m = re.findall(r"scf [a-z ]+:\s+(\d+) \|", buf.getvalue())
    out = {k: getattr(mol, k) for k in want}
    out["notconverged"] = es.notconverged
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    out["n_scf_iter"] = int(m[-1]) if m else -1
    return out, dt
