"""ctypes binding of libseqm_b200.so (include/seqm_b200.h).

The product loads the CUDA library only and refuses to work without a CUDA device: there is no CPU
fallback.  (tests/ may bind the host-emulation build of the same kernel sources through
`SeqmLib(path)` explicitly to check kernel logic on GPU-less CI; nothing in this package does.)
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEQM_B200_LIB") or os.path.join(_HERE, "lib", "libseqm_b200.so")  # env: kernel experiments
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--shared",
              "-Xcompiler", "-fPIC"]  # fmt: skip

# rows of the per-atom parameter table (enum seqm_par_row)
PAR_ROWS = ["U_ss", "U_pp", "zeta_s", "zeta_p", "beta_s", "beta_p", "g_ss", "g_sp", "g_pp", "g_p2", "h_sp", "alpha",
            "Gaussian1_K", "Gaussian2_K", "Gaussian3_K", "Gaussian4_K", "Gaussian1_L", "Gaussian2_L", "Gaussian3_L",
            "Gaussian4_L", "Gaussian1_M", "Gaussian2_M", "Gaussian3_M", "Gaussian4_M", "tore", "qn", "rho_core", "atomic_num",
            "dd", "qq", "rho0", "rho1", "rho2",
            # PM6 d-shell rows (per element, pyseqm_b200/seqm_functions/pm6d_tables.py)
            "U_dd", "zeta_d", "beta_d", "qnd", "dp", "ds", "ddq", "rho3", "rho4", "rho5", "rho6", "rho2d"]  # fmt: skip
NPAR = len(PAR_ROWS)
N_ELEM_ROWS = 28  # rows 0..27 come from the element table; 28..32 are written by seqm_atom_multipoles
METHOD_ID = {"MNDO": 0, "AM1": 1, "PM3": 2, "PM6_SP": 3, "PM6_D": 4}


class SeqmBatchStruct(C.Structure):
    _fields_ = [
        ("nmol", C.c_int32), ("nat", C.c_int32), ("npairs", C.c_int32), ("method", C.c_int32),
        ("nmax", C.c_int32), ("molsize", C.c_int32), ("mat_total", C.c_int64),
        ("mol_atom0", C.c_void_p), ("mol_pair0", C.c_void_p), ("mol_mat0", C.c_void_p),
        ("mol_nheavy", C.c_void_p), ("mol_nhyd", C.c_void_p), ("mol_nocc", C.c_void_p), ("mol_order", C.c_void_p),
        ("atom_Z", C.c_void_p), ("atom_mol", C.c_void_p), ("pair_i", C.c_void_p), ("pair_j", C.c_void_p),
        ("atom_par", C.c_void_p), ("cls_begin", C.c_int32 * 12), ("cls_count", C.c_int32 * 12),
        ("pw_alpha", C.c_void_p), ("pw_chi", C.c_void_p), ("pw_dim", C.c_int32),
        ("pair_cls_off", C.c_int32 * 4), ("pair_perm", C.c_void_p), ("fock_scratch", C.c_int32),
        ("pair_outer_cutoff", C.c_double),
        ("mol_nsh", C.c_void_p), ("pair_wd0", C.c_void_p), ("ypairs", C.c_void_p), ("ypair_slot", C.c_void_p),
        ("n_ypairs", C.c_int32), ("oc_dim", C.c_int32), ("onecenter_d", C.c_void_p), ("mp_coef", C.c_void_p),
        ("mp_coef_yx", C.c_void_p), ("ovl_poly", C.c_void_p), ("wd", C.c_void_p), ("hab_d", C.c_void_p),
    ]  # fmt: skip

JACOBI_NP = (4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64)


class SeqmScfOpts(C.Structure):
    _fields_ = [("eps", C.c_double), ("converger", C.c_int32), ("alpha", C.c_double), ("use_sp2", C.c_int32),
                ("sp2_eps", C.c_double), ("max_iter", C.c_int32), ("warm_start", C.c_int32),
                ("pipeline", C.c_int32)]  # fmt: skip


class SeqmPlanCounts(C.Structure):  # mirrors seqm_plan_counts_t
    _fields_ = [("nat", C.c_int32), ("npairs", C.c_int32), ("nmax", C.c_int32), ("zmax", C.c_int32),
                ("mat_total", C.c_int64), ("odd_electrons", C.c_int32), ("unsorted", C.c_int32),
                ("pairs_overflow", C.c_int32), ("fock_scratch", C.c_int32), ("pair_cls_cnt", C.c_int32 * 3),
                ("jacobi_cls_cnt", C.c_int32 * 13), ("elements", C.c_int32 * 128)]  # fmt: skip


class SeqmError(RuntimeError):
    pass


SOURCES = ("seqm_b200.cu", "seqm_pair.cu", "seqm_spd.cu", "seqm_eigh.cu", "seqm_post.cu", "seqm_ksa.cu")  # translation units, compiled in parallel
# files that only one translation unit includes (everything else is shared): an edit there recompiles that unit alone
_ONLY = {"seqm_spd.cu": {"seqm_spd.cu", "spd_kernels.cuh"},
         "seqm_eigh.cu": {"seqm_eigh.cu", "hestenes_kernels.cuh"},
         "seqm_post.cu": {"seqm_post.cu"},
         "seqm_ksa.cu": {"seqm_ksa.cu"},
         "seqm_pair.cu": {"seqm_pair.cu", "pair_kernels.cuh"},  # ~25 min of nvcc: keep edits out of it
         "seqm_b200.cu": {"seqm_b200.cu", "atom_kernels.cuh", "scf_driver.cuh", "plan_kernels.cuh", "eig_kernels.cuh",
                          "fock_kernels.cuh", "large_kernels.cuh"}}  # fmt: skip


def build_library(verbose=False):
    """Compile pyseqm_b200/csrc for sm_100a into pyseqm_b200/lib/libseqm_b200.so (nvcc cross-compiles without a GPU)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))
    hdr = os.path.join(_HERE, "..", "include", "seqm_b200.h")
    newest = max(newest, os.path.getmtime(hdr))
    stamp = LIB_PATH + ".sources"  # the translation units the existing library was linked from
    linked = open(stamp).read().split() if os.path.exists(stamp) else []
    objs_all = [os.path.join(os.path.dirname(LIB_PATH), src.replace(".cu", ".o")) for src in SOURCES]
    have_objs = all(os.path.exists(o) for o in objs_all)
    if os.path.exists(LIB_PATH) and linked == list(SOURCES) and not have_objs and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH  # a shipped library without its objects (nothing to compare against but the sources)
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(os.path.dirname(LIB_PATH), src.replace(".cu", ".o"))
        objs.append(obj)
        others = set().union(*(v for k, v in _ONLY.items() if k != src))
        deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f not in others] + [hdr]
        if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(d) for d in deps):
            continue
        cmd = ["nvcc"] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if (not procs and os.path.exists(LIB_PATH) and linked == list(SOURCES)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(o) for o in objs)):
        return LIB_PATH  # every object is newer than its sources and the library is newer than every object
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as f:
        f.write(" ".join(SOURCES) + "\n")
    return LIB_PATH


class SeqmLib:
    """Thin typed wrapper around the shared library."""

    _P = C.c_void_p

    def __init__(self, path):
        if not os.path.exists(path):
            raise SeqmError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
        self.path = path
        self.dll = C.CDLL(path)
        P, B, O = self._P, C.POINTER(SeqmBatchStruct), C.POINTER(SeqmScfOpts)
        sig = {
            "seqm_abi_version": ([], C.c_int),
            "seqm_last_error": ([], C.c_char_p),
            "seqm_max_orbitals": ([], C.c_int),
            "seqm_max_orbitals_eig": ([], C.c_int),
            "seqm_plan_count": ([P, C.c_int32, C.c_int32, P, P, C.c_int32, P, P, P, P, P, P, P, P, P,
                                 C.POINTER(SeqmPlanCounts), P, P], C.c_int),
            "seqm_plan_fill": ([P, C.c_int32, C.c_int32, C.POINTER(SeqmPlanCounts), P, P, P, P, P, C.c_int32, C.c_int32,
                                P, P, P, P, P, P, P, P], C.c_int),
            "seqm_atom_multipoles": ([B, P], C.c_int),
            "seqm_pair_integrals": ([B, P, P, P, P], C.c_int),
            "seqm_pair_integrals_d": ([B, P, P, P, P, P], C.c_int),
            "seqm_hcore": ([B, P, P, P, P], C.c_int),
            "seqm_fock": ([B, P, P, P, P, P, P], C.c_int),
            "seqm_eig_density": ([B, P, P, P, P, P, P, P], C.c_int),
            "seqm_sp2_density": ([B, P, P, C.c_double, P, P, P], C.c_int),
            "seqm_sp2_large_workspace_bytes": ([B], C.c_int64),
            "seqm_sp2_density_large": ([B, P, P, C.c_double, C.POINTER(C.c_int32), P, P], C.c_int),
            "seqm_elec_energy": ([B, P, P, P, P, P, P], C.c_int),
            "seqm_nuclear_energy": ([B, P, P, P, P, P], C.c_int),
            "seqm_gradient": ([B, P, P, P, P, P], C.c_int),
            "seqm_pack": ([B, P, P, P], C.c_int),
            "seqm_unpack": ([B, P, P, P], C.c_int),
            "seqm_initial_density": ([B, P, P], C.c_int),
            "seqm_scf_workspace_bytes": ([B, O], C.c_int64),
            "seqm_scf": ([B, O, P, P, P, P, P, P, P, C.POINTER(C.c_int32), P, P], C.c_int),
            "seqm_gradient_forward": ([B, P, P, P, P, P], C.c_int),
            "seqm_gradient_xl": ([B, P, P, P, P, P, P], C.c_int),
            "seqm_elec_energy_xl": ([B, P, P, P, P, P, P], C.c_int),
            "seqm_xl_propagate": ([C.c_int64, C.c_double, C.c_double, P, P, P, P, C.c_int32, C.c_int32, P, P], C.c_int),
            "seqm_orbitals_dense": ([B, P, P, P], C.c_int),
            "seqm_post_scf": ([B, P, P, P, P, P, P, C.c_double, C.c_double, P], C.c_int),
            "seqm_mo_match": ([B, P, P, P, P, P, P, P, P, P, P], C.c_int),
            "seqm_launch_count": ([], C.c_longlong),
            "seqm_fp64_peak_tflops": ([], C.c_double),
            "seqm_square_product": ([C.c_int, P, P, P, P], C.c_int),
            "seqm_packed_gemm": ([B, P, P, P, C.c_int, C.c_int, P], C.c_int),
            "seqm_scale_columns": ([B, P, P, C.c_double, P, P], C.c_int),
            "seqm_canon_prt": ([B, P, P, P, C.c_double, C.c_int, P], C.c_int),
            "seqm_packed_dot": ([B, P, P, P, P], C.c_int),
            "seqm_packed_axpby": ([B, P, P, P, P, P], C.c_int),
            "seqm_fock_antisym": ([B, P, P, P, P], C.c_int),
            "seqm_packed_transpose": ([B, P, P, P], C.c_int),
            "seqm_jacobi_stats": ([C.POINTER(C.c_ulonglong), C.c_int], C.c_int),
            "seqm_profile_enable": ([C.c_int], C.c_int),
            "seqm_profile_kinds": ([], C.c_int),
            "seqm_profile_name": ([C.c_int], C.c_char_p),
            "seqm_profile_collect": ([C.POINTER(C.c_double), C.POINTER(C.c_int32)], C.c_int),
        }
        for name, (args, res) in sig.items():
            fn = getattr(self.dll, name)
            fn.argtypes = args
            fn.restype = res
        self.symbols = list(sig)
        if self.dll.seqm_abi_version() != 3:
            raise SeqmError("libseqm_b200 ABI version mismatch")

    def jacobi_stats(self, reset=True):
        out = (C.c_ulonglong * 8)()
        self.check(self.dll.seqm_jacobi_stats(out, 1 if reset else 0), "seqm_jacobi_stats")
        return {"molecules": out[0], "sweeps": out[1], "first_order_finishes": out[2], "no_sweep": out[3],
                "cycles_transform": out[4], "cycles_sweeps": out[5], "cycles_epilogue": out[6], "cycles_total": out[7]}

    def profile_enable(self, on=True):
        self.dll.seqm_profile_enable(1 if on else 0)

    def profile_collect(self):
        """{kernel kind: (milliseconds, launches)} since profile_enable(True)."""
        n = self.dll.seqm_profile_kinds()
        ms = (C.c_double * n)()
        cnt = (C.c_int32 * n)()
        self.check(self.dll.seqm_profile_collect(ms, cnt), "seqm_profile_collect")
        return {self.dll.seqm_profile_name(k).decode(): (ms[k], cnt[k]) for k in range(n)}

    def check(self, rc, what):
        if rc != 0:
            raise SeqmError(f"{what} failed (code {rc}): {self.dll.seqm_last_error().decode()}")


_LIB = None


def get_lib():
    """The CUDA library.  Raises if CUDA is unavailable: pyseqm_b200 has no CPU path."""
    global _LIB
    if _LIB is None:
        if not torch.cuda.is_available():
            raise SeqmError("pyseqm_b200 needs a CUDA device (B200, sm_100a); it has no CPU fallback")
        _LIB = SeqmLib(LIB_PATH)
    return _LIB


def ptr(t):
    if t is None:
        return None
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def stream_of(t):
    if t.is_cuda:
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return C.c_void_p(0)
